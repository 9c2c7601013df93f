"""TEST INFRASTRUCTURE ONLY: compile oracle/qoracle.c -> oracle/libqoracle.so with gcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "qoracle.c")
SO = os.path.join(HERE, "libqoracle.so")


def build(force=False):
    if (not force) and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fopenmp", SRC, "-o", SO, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force=True))
