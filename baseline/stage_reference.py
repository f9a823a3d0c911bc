"""Copy the reference's Python tree into baseline/_ref/R-PCC (git-ignored).  Byte-for-byte copies, nothing edited;
the native code it imports comes from oracle/_ref (reference sources compiled where they lie) or from rpcc_b200.plugin."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = (".py", ".yaml", ".yml", ".txt", ".csv", ".md")


def stage(ref="/root/reference", dst=None):
    dst = dst or os.path.join(HERE, "_ref", "R-PCC")
    if not os.path.isdir(os.path.join(ref, "tools")):
        return None
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for top in ("tools", "utils", "dataset", "cfgs", "data"):
        src = os.path.join(ref, top)
        if not os.path.isdir(src):
            continue
        for dp, dn, fs in os.walk(src):
            dn[:] = [d for d in dn if d not in (".git", "__pycache__")]
            for f in fs:
                if f.endswith(KEEP):
                    out = os.path.join(dst, os.path.relpath(os.path.join(dp, f), ref))
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    shutil.copyfile(os.path.join(dp, f), out)
    os.makedirs(os.path.join(dst, "ops", "fps"), exist_ok=True)
    shutil.copyfile(os.path.join(ref, "ops", "fps", "fps_utils.py"), os.path.join(dst, "ops", "fps", "fps_utils.py"))
    ex = os.path.join(ref, "assets", "example_data", "example.bin")
    if os.path.exists(ex):
        os.makedirs(os.path.join(dst, "assets", "example_data"), exist_ok=True)
        shutil.copyfile(ex, os.path.join(dst, "assets", "example_data", "example.bin"))
    return dst


if __name__ == "__main__":
    print(stage(*sys.argv[1:]))
