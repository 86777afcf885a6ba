"""Build the CPU oracle (test infrastructure) into oracle/libgenoracle.so.

The reference (Gen.jl) is pure Julia and Julia is absent from this image, so there is
no oracle/_ref: nothing of the reference can be compiled here ("Oracle").
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gen_oracle.c")
OUT = os.path.join(HERE, "libgenoracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(SRC[:-2] + ".h"),
                                             os.path.getmtime(os.path.join(HERE, "wexp_table.inc")))):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=gnu11",
           "-Wall", "-Wextra", "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
