"""Stage-by-stage GPU diagnostic against the CPU oracle (development tool, run under gpurun).
Prints one block per stage and keeps going after a failure so one GPU call exposes as much as possible."""
import sys, time, traceback
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import ngicp, oracle
from oracle import voxel_keys as vk
from ngicp import synth
import scenarios as S

BIG = "--big" in sys.argv
def stage(name):
    print(f"\n===== {name} =====", flush=True)

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

def run(name, fn):
    stage(name)
    t0 = time.time()
    try:
        fn()
        print(f"[{name}] done in {time.time()-t0:.2f}s", flush=True)
    except Exception:
        traceback.print_exc()
        print(f"[{name}] FAILED", flush=True)

a, b, T_true = S.scan_pair(0, w=128)
print("clouds", a.shape, b.shape)
g = ngicp.NanoGICP(0)
S.configure(g)
og = S.configure(oracle.OracleGICP("port"))
state = {}

def st_keys():
    g.setInputSource(a)
    keys, lo, h0 = g.source_kdtree_.voxel_keys()
    lo_o, h0_o = vk.grid_params(a)
    print("origin", lo, lo_o, "h0", h0, h0_o)
    ko = vk.voxel_keys(a, lo_o, h0_o)
    print("keys equal:", bool((keys == ko).all()), "mismatch", int((keys != ko).sum()), "of", len(keys))
run("K1 keys", st_keys)

def st_knn():
    tree = g.source_kdtree_
    idx, sqd = tree.nearestKSearch(a, 16)
    ot = oracle.KdTree(a, "port")
    oi, od = ot.knn(a, 16)
    print("rows exact/tie/bad:", S.knn_rows_equivalent(idx, sqd, oi, od))
    print("dist bit-equal:", bool((sqd == od).all()), " idx equal frac:", float((idx == oi).all(1).mean()))
    bad = np.nonzero(~(idx == oi).all(1))[0][:3]
    for r in bad:
        print(" row", r, "\n  gpu", idx[r], sqd[r], "\n  orc", oi[r], od[r])
    for k in (1, 5, 20, 32, 40):
        idx, sqd = tree.nearestKSearch(b[:2000], k)
        oi, od = ot.knn(b[:2000], k)
        print(f" k={k} cross-query rows exact/tie/bad:", S.knn_rows_equivalent(idx, sqd, oi, od))
    far = (a[:64] + np.float32([500, -300, 40])).astype(np.float32)
    idx, sqd = tree.nearestKSearch(far, 4); oi, od = ot.knn(far, 4)
    print(" far outside-grid queries exact/tie/bad:", S.knn_rows_equivalent(idx, sqd, oi, od))
run("K2 knn", st_knn)

def st_cov():
    g.calculateSourceCovariances()
    C = g.getSourceCovariances()
    og.setInputSource(a); og.calculateSourceCovariances(); Co = og.getSourceCovariances()
    oi, _ = oracle.KdTree(a, "port").knn(a, 16)
    ok = S.spectral_gap_ok(a, oi)
    err = np.abs(C - Co).reshape(len(C), -1).max(1)
    print("density gpu/oracle", g.source_density_, og.source_density_)
    print("gap-ok rows", int(ok.sum()), "/", len(ok), " max abs err (ok rows)", float(err[ok].max()), " all rows", float(err.max()),
          " rows>1e-4:", int((err[ok] > 1e-4).sum()))
    print("NaNs:", int(np.isnan(C).sum()))
    for reg in (ngicp.REG_NONE, ngicp.REG_MIN_EIG, ngicp.REG_NORMALIZED_MIN_EIG, ngicp.REG_FROBENIUS):
        g.setRegularizationMethod(reg); og.setRegularizationMethod(reg)
        g.calculateSourceCovariances(); og.calculateSourceCovariances()
        C2, Co2 = g.getSourceCovariances(), og.getSourceCovariances()
        e2 = np.abs(C2 - Co2).reshape(len(C2), -1).max(1) / np.maximum(np.abs(Co2).reshape(len(C2), -1).max(1), 1e-12)
        print(f" reg {reg}: max rel err ok-rows {float(e2[ok].max()):.3e} all {float(e2.max()):.3e}")
    g.setRegularizationMethod(ngicp.REG_PLANE); og.setRegularizationMethod(ngicp.REG_PLANE)
    g.calculateSourceCovariances(); og.calculateSourceCovariances()
run("K3 covariances", st_cov)

def st_lin():
    g.setInputTarget(b); g.calculateTargetCovariances()
    og.setInputTarget(b); og.calculateTargetCovariances()
    # swap roles: align a (source) onto b (target)
    for name, T in (("identity", np.eye(4)), ("perturbed", synth.se3((0.01, -0.02, 0.015), (0.2, -0.1, 0.05)))):
        e, H, bb = g.linearize(T)
        eo, Ho, bo = og.linearize(T)
        print(f"[{name}] err {e:.9g} vs {eo:.9g} rel {abs(e-eo)/abs(eo):.2e}; H rel {rel(H,Ho):.2e}; b rel {rel(bb,bo):.2e}; ncorr {g.num_correspondences} vs {og.num_correspondences}")
        T2 = synth.se3((0.001, 0.002, -0.001), (0.01, 0.02, -0.01)) @ T
        e2, eo2 = g.compute_error(T2), og.compute_error(T2)
        print(f"   compute_error {e2:.9g} vs {eo2:.9g} rel {abs(e2-eo2)/abs(eo2):.2e}")
        corr, sqd, mah = g.update_correspondences(T)
        co, so, mo = og.update_correspondences(T)
        both = (corr >= 0) & (co >= 0)
        print(f"   corr equal {int((corr==co).sum())}/{len(corr)}  sqd bit-equal on valid {bool((sqd[both]==so[both]).all())}  mahal rel {rel(mah[both], mo[both]):.2e}")
run("K4/K5 linearize + error", st_lin)

def st_align():
    T = g.align()
    To = og.align()
    print("gpu iters", g.nr_iterations_, g.converged_, " oracle iters", og.nr_iterations_, og.converged_)
    print("dT max", float(np.abs(T - To).max()), " trans", float(np.abs(T[:3,3]-To[:3,3]).max()))
    print(T); print(To); print("truth (b->a inverse)"); print(np.linalg.inv(T_true))
    print("final err", g.getFinalError(), og.getFinalError())
run("align", st_align)

def st_batch():
    sc = synth.Scene(3); rng = np.random.default_rng(5)
    poses = synth.trajectory(sc, 4, 3)
    clouds = [synth.voxel_filter(synth.scan(sc, P, rng, w=96)) for P in poses]
    pts = np.concatenate(clouds); off = np.cumsum([0] + [len(c) for c in clouds])
    gg = S.configure(ngicp.NanoGICP(0))
    cov6, m4, dens = gg.batchCovariances(pts, off, want_mat4=True)
    worst = 0; 
    for s, c in enumerate(clouds):
        o = S.configure(oracle.OracleGICP("port")); o.setInputSource(c); o.calculateSourceCovariances()
        Co = o.getSourceCovariances(); oi, _ = oracle.KdTree(c, "port").knn(c, 16); ok = S.spectral_gap_ok(c, oi)
        e = np.abs(m4[off[s]:off[s+1]] - Co).reshape(len(c), -1).max(1)
        worst = max(worst, float(e[ok].max()))
        print(f" seg {s}: n={len(c)} max err ok rows {float(e[ok].max()):.3e} density {dens[s]:.6g} vs {o.source_density_:.6g}")
    print("batch worst", worst)
run("batch covariances", st_batch)

def st_big():
    sc = synth.Scene(0); rng = np.random.default_rng(2)
    t0 = time.time()
    tgt, bounds, poses = synth.make_submap(sc, 1_000_000, 0, n_keyframes=40)
    T_ws = poses[20] @ synth.se3((0, 0, 0.02), (0.3, 0.1, 0.0))
    src = synth.transform_points(T_ws, synth.scan(sc, T_ws, rng, keep_all=True))
    T_off = synth.se3((0.01, -0.015, 0.03), (0.15, -0.1, 0.05))
    src = synth.transform_points(np.linalg.inv(T_off), src)
    print("data gen", time.time() - t0, src.shape, tgt.shape)
    G = S.configure(ngicp.NanoGICP(0)); G.enableTiming(True)
    t0 = time.time(); G.setInputTarget(tgt); G.synchronize(); print("target index wall", time.time() - t0, G.timings())
    t0 = time.time(); G.calculateTargetCovariances(); print("target cov wall", time.time() - t0, G.timings())
    for it in range(3):
        t0 = time.time(); G.setInputSource(src); G.calculateSourceCovariances(); t1 = time.time()
        T = G.align(); t2 = time.time()
        print(f"run {it}: source prep {1e3*(t1-t0):.2f} ms, align {1e3*(t2-t1):.2f} ms, iters {G.nr_iterations_} conv {G.converged_}", G.timings())
    G.enableTiming(False)
    for it in range(3):
        src2 = src.copy()
        t0 = time.time(); G.setInputSource(src2); G.calculateSourceCovariances(); t1 = time.time()
        T = G.align(); t2 = time.time()
        print(f"untimed run {it}: source prep {1e3*(t1-t0):.2f} ms, align {1e3*(t2-t1):.2f} ms, launches {G.timings()['kernel_launches']}")
    print("recovered vs T_off max diff", float(np.abs(T - T_off).max())); print(T); print(T_off)
    O = S.configure(oracle.OracleGICP("port"))
    t0 = time.time(); O.setInputTarget(tgt); O.calculateTargetCovariances(); print("oracle target prep", time.time() - t0)
    t0 = time.time(); O.setInputSource(src); O.calculateSourceCovariances(); t1 = time.time(); To = O.align(); t2 = time.time()
    print(f"oracle source prep {t1-t0:.3f}s align {t2-t1:.3f}s iters {O.nr_iterations_} conv {O.converged_}")
    print("gpu vs oracle dT", float(np.abs(T - To).max()), "trans", float(np.abs(T[:3,3]-To[:3,3]).max()))
if BIG:
    run("big: 65k scan vs 1M submap", st_big)
print("\nDIAG COMPLETE")
