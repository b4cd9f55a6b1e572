"""Pack the public eBOSS DR16 NGC inputs the configs of BASELINE.json name (window functions, data
multipoles, covariances; /root/reference/data/DR16_noric, described in its README.md) into one
compact fixture, `eftpipe_b200/data/dr16_ngc.npz`, so that tests and bench.py can run on the GPU box
where /root/reference does not exist.  Run once in the build container; output is committed."""
import os
import sys

import numpy as np

SRC = os.path.join(os.environ.get("EFTPIPE_REFERENCE", "/root/reference"), "data", "DR16_noric")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "eftpipe_b200", "data", "dr16_ngc.npz")


def header_value(path, key):
    with open(path) as fh:
        for line in fh:
            if line.startswith("#") and key in line:
                return float(line.split("=")[1])
    return np.nan


def main():
    out = {}
    for t in ("LRG", "ELG", "X"):
        w = np.loadtxt(os.path.join(SRC, f"win_NGC_{t}.txt"))
        out[f"win_{t}"] = np.ascontiguousarray(w[:, :4])  # s, Q0, Q2, Q4 (Window default Nq=3)
    for name in ("NGC_LRG_P", "NGC_ELG_Q", "NGC_X_P", "NGC_ELG_P"):
        path = os.path.join(SRC, name + ".txt")
        out[name] = np.loadtxt(path)
        out[name + "_Pshot"] = header_value(path, "Pshot")
    for name in ("cov_NGC_L024E02X024_PQP", "cov_NGC_L024_P"):
        out[name] = np.loadtxt(os.path.join(SRC, name + ".txt"))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
