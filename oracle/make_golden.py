"""TEST INFRASTRUCTURE ONLY: mints the golden vectors under tests/golden/ from the reference itself.

The reference ships no tests, fixtures or known-answer vectors, so the pins
are minted here by running the reference's OWN unmodified sources (compiled into oracle/_ref by
oracle/Makefile) on small synthetic inputs, and committing inputs + outputs:

  tests/golden/events.npz   the input slices (fr_x, fr_y, t_ns per case; small)
  tests/golden/golden.json  per case: setup, iteration count, dividers, the 11 model scalars as
                            hex floats, SHA-256 of the per-event outputs, and stage-level records
                            (iteration-0 time image / Scharr images / model of that image)

Run in the build container (needs /root/reference):   python oracle/make_golden.py
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from better_flow_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hexes(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64)]


def minimise_case(name, sl, rows, cols, scale, max_iter, init=None, noise=None):
    r = ref.minimize(sl["fr_x"], sl["fr_y"], sl["t_ns"], scale=scale, max_iter=max_iter, init_model=init,
                     noise=noise, rows=rows, cols=cols, want_events=True)
    return {
        "name": name, "events": sl["key"], "rows": rows, "cols": cols, "scale": scale, "max_iter": max_iter,
        "init": hexes(init) if init is not None else None,
        "noise_every": None,
        "rc": r["rc"], "iters": r["iters"], "model": hexes(r["model"]),
        "dividers": [float(v) for v in r["dividers"]],
        "setup": {k: r[k] for k in ("x_min", "x_max", "y_min", "y_max", "wsize_x", "wsize_y", "img_rows", "img_cols",
                                     "x_shift", "y_shift")},
        "sha_pr_x": sha(r["pr_x"]), "sha_pr_y": sha(r["pr_y"]), "sha_nx": sha(r["nx"]), "sha_ny": sha(r["ny"]),
        "pr_x_head": hexes(r["pr_x"][:4]), "nx_head": hexes(r["nx"][:4]),
        "noise_out_sum": int(r["noise"].sum()),
    }, r


def stage_positions(fr_x, fr_y):
    """Deterministic sub-pixel jitter (no RNG, exactly reproducible): some events leave the window."""
    i = np.arange(len(fr_x), dtype=np.int64)
    pr_x = fr_x.astype(np.float64) + ((i * 7919) % 1000 - 500).astype(np.float64) / 256.0
    pr_y = fr_y.astype(np.float64) + ((i * 104729) % 1000 - 500).astype(np.float64) / 256.0
    return pr_x, pr_y


def main():
    os.makedirs(OUT, exist_ok=True)
    events = {}
    cases, stages = [], []

    def keep(key, s):
        events[key + "_fr_x"] = s.fr_x.astype(np.uint16)
        events[key + "_fr_y"] = s.fr_y.astype(np.uint16)
        events[key + "_t_ns"] = s.t_ns.astype(np.int32)
        return {"key": key, "fr_x": s.fr_x, "fr_y": s.fr_y, "t_ns": s.t_ns}

    # --- DAVIS-240C, translating contours: BASELINE configs[0] / [1] shaped ---------------------------
    st = synth.make_stream(240, 180, 3e6, 0.02, seed=1)
    a, b = synth.cut_slices(st, 0.01)[:2]
    A, B = keep("davis240_a", a), keep("davis240_b", b)
    c, ra = minimise_case("davis240_10ms_maxiter10", A, 180, 240, 3, 10); cases.append(c)
    c, _ = minimise_case("davis240_10ms_converge", A, 180, 240, 3, -1); cases.append(c)
    c, _ = minimise_case("davis240_10ms_scale1_maxiter25", A, 180, 240, 1, 25); cases.append(c)
    c, _ = minimise_case("davis240_10ms_scale5_maxiter6", A, 180, 240, 5, 6); cases.append(c)
    c, _ = minimise_case("davis240_warmstart", B, 180, 240, 3, 10, init=ra["model"]); cases.append(c)
    # pre-marked noise events
    noise = (np.arange(len(a.fr_x)) % 7 == 0).astype(np.uint8)
    c, _ = minimise_case("davis240_noise_every7", A, 180, 240, 3, 5, noise=noise); c["noise_every"] = 7; cases.append(c)

    # --- rotating + expanding scene (exercises rot / div) --------------------------------------------
    st = synth.make_stream(240, 180, 3e6, 0.01, seed=21, vel=(-150.0, 60.0), omega=1.5, expand=0.8)
    R = keep("davis240_rot", synth.cut_slices(st, 0.01)[0])
    c, _ = minimise_case("davis240_rotating_maxiter40", R, 180, 240, 3, 40); cases.append(c)

    # --- guards ---------------------------------------------------------------------------------------
    few = {"key": "davis240_a", "fr_x": a.fr_x[:999], "fr_y": a.fr_y[:999], "t_ns": a.t_ns[:999]}
    c, _ = minimise_case("guard_fewer_than_1000", few, 180, 240, 3, -1); c["first_n"] = 999; cases.append(c)

    # --- DAVIS-346, 50 ms slice (configs[2] shaped) -------------------------------------------------
    if ref.available(260, 346):
        st = synth.make_stream(346, 260, 2e6, 0.05, seed=3)
        S = keep("davis346", synth.cut_slices(st, 0.05)[0])
        c, _ = minimise_case("davis346_50ms_maxiter10", S, 260, 346, 3, 10); cases.append(c)

    # --- stage-level: iteration-0 image of case A with jittered positions -----------------------------
    pr_x, pr_y = stage_positions(a.fr_x, a.fr_y)
    for scale in (1, 3, 5):
        su = ref.minimize(a.fr_x, a.fr_y, a.t_ns, scale=scale, max_iter=1)
        img = ref.time_img(pr_x, pr_y, a.t_ns, su["wsize_x"], su["wsize_y"], scale, int(su["x_shift"]), int(su["y_shift"]))
        m7, gx, gy = ref.model(img, want_grad=True)
        stages.append({"scale": scale, "w": su["wsize_x"], "h": su["wsize_y"], "x_sh": int(su["x_shift"]),
                       "y_sh": int(su["y_shift"]), "sha_img": sha(img), "img_sum": float(img.astype(np.float64).sum()).hex(),
                       "nnz": int((img > 0).sum()), "model7": hexes(m7), "sha_gx": sha(gx), "sha_gy": sha(gy)})
    # projection KAT
    args = (-0.043, 0.081, 91.3, 118.7, 3.1e-5, -2.2e-4)
    px, py, nx, ny = ref.project(a.fr_x, a.fr_y, a.t_ns, pr_x, pr_y, *args)
    proj = {"args": hexes(args), "sha_pr_x": sha(px), "sha_pr_y": sha(py), "sha_nx": sha(nx), "sha_ny": sha(ny)}

    # --- whole-stream DVS_flow (ring buffer, triggers, warm start): reference CLI configuration -------
    st = synth.make_stream(240, 180, 1e6, 0.075, seed=9)
    events["stream_x"] = st.x
    events["stream_y"] = st.y
    events["stream_t_ns"] = st.t_ns.astype(np.int32)
    streams = []
    for stm in (False, True):
        models, info = ref.stream(st.y, st.x, st.t_ns, config=0, ev_refresh=20000, time_refresh_ns=33000000,
                                  scale=3, max_iter=10, stm_disable=stm, flush=True)
        streams.append({"config": 0, "ev_refresh": 20000, "time_refresh_ns": 33000000, "scale": 3, "max_iter": 10,
                        "stm_disable": stm, "n_slices": int(len(models)),
                        "models": [hexes(m) for m in models], "info": info.tolist()})

    np.savez_compressed(os.path.join(OUT, "events.npz"), **events)
    json.dump({"minted_from": "oracle/_ref (reference sources compiled unmodified)", "cases": cases, "stages": stages,
               "project": proj, "streams": streams, "event_sha": {k: sha(v) for k, v in events.items()}},
              open(os.path.join(OUT, "golden.json"), "w"), indent=1)
    print("wrote", len(cases), "cases,", len(stages), "stage records,", len(streams), "streams;",
          os.path.getsize(os.path.join(OUT, "events.npz")) // 1024, "KiB of events")


if __name__ == "__main__":
    main()
