"""Build the C restatement of the greedy step (oracle/greedy_step.c) -> oracle/_build/libgreedy_oracle.so.
TEST INFRASTRUCTURE ONLY.  -ffp-contract=off keeps the arithmetic identical to the NumPy oracle (no FMA)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libgreedy_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "greedy_step.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
           src, "-o", OUT, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
