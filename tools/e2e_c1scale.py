"""C1-scale end-to-end check through the REAL ``WGCNA.findModules`` (PyWGCNA/wgcna.py:265-404) with
``pywgcna_b200.install()`` on the reference class: the 192 x 23 844 stand-in for the 5xFAD tutorial matrix
(seed 5192), networkType signed hybrid / unsigned, TOM arithmetic i8 / f64.  The three hot-path static
methods run on the GPU, everything after them (round(1 - TOM, 8), squareform, linkage, cutreeHybrid, labels2colors,
moduleEigengenes, mergeCloseModules) is the reference's own host code out of baseline/_ref; the outcome is compared
with the fixtures the UNMODIFIED reference produced in the build container (tests/golden/c1scale_*.npz,
oracle/gen_golden_large.py): selected power, soft-threshold table, gene tree merge order, cutreeHybrid labels,
merged module labels.

The host part takes minutes per variant (cutreeHybrid alone ~4 min), so the variants run as parallel processes
sharing the GPU.   python tools/e2e_c1scale.py [out.json] [variant ...]     variant = "<net>:<prec>"
"""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200WGCNA_QUIET", "1")
os.environ.setdefault("B200W_REF", os.path.join(ROOT, "baseline", "_ref"))

VARIANTS = ["signed hybrid:i8", "signed hybrid:f64", "unsigned:i8", "unsigned:f64"]


def run_variant(variant):
    import numpy as np

    import pywgcna_b200 as bw
    from oracle import findmodules_harness as H
    from oracle import ref_import as R
    from pywgcna_b200.synth import CONFIGS, make_expression

    net, prec = variant.split(":")
    tom = "signed" if net == "signed hybrid" else "unsigned"
    g = np.load(os.path.join(ROOT, "tests", "golden", f"c1scale_{net.replace(' ', '_')}.npz"))
    m, n, F, seed, _, _ = CONFIGS["C1"]
    X = make_expression(m, n, F=F, seed=seed)
    W = R.load_reference()
    bw.wgcna.PRECISION = prec
    bw.wgcna.VERBOSE = False
    bw.install(W.WGCNA)  # the real class: findModules looks the three methods up on it by name (:278, :310, :320)
    assert W.WGCNA.__dict__["TOMsimilarity"].__func__ is bw.TOMsimilarity
    assert W.linkage is bw.linkage  # wgcna.py:328 runs the GPU average linkage as well
    lib = bw._lib.load()
    launches0 = lib.b200w_launch_count()
    obj = H.make_wgcna(X, networkType=net, TOMType=tom)
    t0 = time.perf_counter()
    r = H.run_findModules(obj)
    wall = time.perf_counter() - t0
    Z = r["linkage"]
    T = obj.TOM.to_numpy()
    rows = g["rows"]
    res = {
        "variant": variant, "genes": n, "samples": m,
        "power": int(r["power"]), "power_ref": int(g["power"]), "power_identical": int(r["power"]) == int(g["power"]),
        "sft_max_abs": float(np.max(np.abs(r["sft"] - g["sft"]))),
        "tom_rows_max_abs": float(np.max(np.abs(T[rows] - g["tom_rows"]))),
        "merge_order_identical": bool(np.array_equal(Z[:, 0].astype(np.int32), g["link_a"])
                                      and np.array_equal(Z[:, 1].astype(np.int32), g["link_b"])
                                      and np.array_equal(Z[:, 3].astype(np.int32), g["link_n"])),
        "merge_height_max_abs": float(np.max(np.abs(Z[:, 2] - g["link_h"]))),
        "dynamic_labels_identical": bool(np.array_equal(r["dynamicMods"], g["dynamicMods"])),
        "dynamic_labels_differing": int((r["dynamicMods"] != g["dynamicMods"]).sum()),
        "module_labels_identical": bool(np.array_equal(r["moduleLabels"], g["moduleLabels"])),
        "modules_dynamic": int(len(np.unique(r["dynamicMods"]))), "modules_merged": int(len(np.unique(r["moduleLabels"]))),
        "gpu_kernels_launched": int(lib.b200w_launch_count() - launches0),
        "seconds": {k: round(v, 3) for k, v in r["timings"].items()}, "wall_s": round(wall, 1),
        "reference_seconds": dict(zip(["pickSoftThreshold", "adjacency", "TOMsimilarity", "cutreeHybrid", "findModules"],
                                      [round(float(x), 1) for x in g["ref_seconds"]])),
    }
    bw.uninstall(W.WGCNA)
    return res


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_e2e_c1scale.json")
    variants = sys.argv[2:] or VARIANTS
    # every variant holds ~25 GB of host memory (adjacency, TOM, dissTOM as DataFrames): as many at a time as fit
    try:
        avail_gb = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) / 1e6
    except Exception:
        avail_gb = 64.0
    try:  # a container's own limit, if it has one
        lim = open("/sys/fs/cgroup/memory.max").read().strip()
        if lim != "max":
            avail_gb = min(avail_gb, int(lim) / 1e9 * 0.8)
    except Exception:
        pass
    procs = max(1, min(len(variants), int(avail_gb // 40)))
    ctx = mp.get_context("spawn")
    results = []

    def report(final):
        ok = all(r["power_identical"] and r["merge_order_identical"] and r["dynamic_labels_identical"]
                 and r["module_labels_identical"] for r in results) and (not final or len(results) == len(variants))
        rep = {"ok": ok, "complete": final, "parallel_processes": procs, "host_mem_available_gb": round(avail_gb, 1),
               "results": sorted(results, key=lambda r: variants.index(r["variant"]))}
        json.dump(rep, open(out, "w"), indent=1)
        return rep

    with ctx.Pool(procs, maxtasksperchild=1) as pool:
        for r in pool.imap_unordered(run_variant, variants, chunksize=1):
            results.append(r)
            report(False)  # partial results survive a timeout
    rep = report(True)
    ok = rep["ok"]
    print(json.dumps(rep))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
