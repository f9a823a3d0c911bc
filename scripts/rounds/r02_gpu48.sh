#!/bin/bash
# round 2, visit 48: the committed tree once more: GPU suite, smoke
exec > gpurun_out/r02n_visit48.txt 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
