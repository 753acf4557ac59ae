"""End-to-end label check in two stages (the reference tree does not travel to the GPU box):

  stage gpu  (on the B200 box):   GPU pickSoftThreshold/adjacency/TOMsimilarity on the e2e fixture input
                                  -> gpurun_out/e2e_tom_<prec>.npy
  stage ref  (build container):   the reference's own linkage + cutreeHybrid (oracle/ref_import.py) on those
                                  TOMs -> labels compared with the golden labels of the unmodified reference
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200WGCNA_QUIET", "1")
import numpy as np, pandas as pd

g = np.load(os.path.join(ROOT, "tests", "golden", "e2e_modules.npz"))
out = os.path.join(ROOT, "gpurun_out")
if sys.argv[1] == "gpu":
    import pywgcna_b200 as bw
    df = pd.DataFrame(g["X"])
    p, _ = bw.pickSoftThreshold(df, networkType="signed hybrid", powerVector=list(g["powers"]))
    A = bw.adjacency(df, adjacencyType="signed hybrid", power=p)
    for prec in ("i8", "i8-fast", "f64"):
        np.save(os.path.join(out, f"e2e_tom_{prec}.npy"), bw.TOMsimilarity(A.copy(), TOMType="signed", precision=prec).to_numpy())
    print("power", int(p), "golden", int(g["power"]))
else:
    from oracle import ref_import as R
    res = {}
    for prec in ("i8", "i8-fast", "f64"):
        T = np.load(os.path.join(out, f"e2e_tom_{prec}.npy"))
        labels, Z = R.ref_modules_from_tom(T, minClusterSize=30)
        res[prec] = {"labels_identical": bool(np.array_equal(labels, g["labels"])), "n_genes": int(len(labels)),
                     "n_modules": int(len(np.unique(labels))), "differing": int((labels != g["labels"]).sum()),
                     "linkage_identical": bool(np.array_equal(Z[:, [0, 1, 3]], g["linkage"][:, [0, 1, 3]]))}
    print(json.dumps(res))
    json.dump(res, open(os.path.join(ROOT, "profiles", "r01_e2e_labels.json"), "w"), indent=1)
