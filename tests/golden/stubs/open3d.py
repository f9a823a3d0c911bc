"""The reference imports open3d at module top (utils/segment_utils.py:7, dataset/dataset.py:3); nothing on
the golden path calls it (the ground model is injected)."""
