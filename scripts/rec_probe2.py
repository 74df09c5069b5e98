import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ.pop("PLAS_DEBUG", None)
import rec_probe
for ng in ("1", "2", "4"):
    os.environ["PLAS_REC_NG"] = ng
    for B, U in ((16, 512), (32, 512), (64, 512), (128, 512), (64, 256)):
        print("NG", ng, end=" ")
        rec_probe.probe(B, U, 400, "tc")
