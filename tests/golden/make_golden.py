#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libg4hepem_ref.so, built by
oracle/Makefile from /root/reference).  Run in a container that has /root/reference:

    python tests/golden/make_golden.py

The vectors pin: VDT log/exp, e-/e+ look-ups (BASELINE configs[0] recipe), gamma cross sections + process choice,
target-element selectors, and full HowFar / Perform / fused steps of small e-/e+ and gamma batches (state in, state
out, secondaries), all with the counter based uniform stream of oracle/g4h_rng_host.h at seed 2026."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from g4hepem_b200 import batches, tables  # noqa: E402
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STATE = os.path.join(HERE, "hepem_state.json")
SEED = 2026


def batch_arrays(b, prefix):
    out = {f"{prefix}{g}": getattr(b, g).copy() for g in b.groups() + ("meta", "winner")}
    return out


def sec_arrays(q, prefix):
    r = q.sorted_records()
    return {f"{prefix}{k}": v for k, v in r.items()}


def main():
    ft = tables.load_state_json(STATE)
    R = ref.Reference(STATE)
    rng = np.random.default_rng(20261017)
    out = {}
    n = 4096
    x = np.exp(rng.uniform(np.log(1e-300), np.log(1e300), n))
    xe = rng.uniform(-720, 720, n)
    out.update(vdt_x=x, vdt_log=R.vdt_log_exp(x)[0], vdt_xe=xe, vdt_exp=R.vdt_log_exp(xe)[1])
    imc = rng.integers(0, ft.num_matcut, n).astype(np.int32)
    ek = np.exp(rng.uniform(np.log(0.95e-4), np.log(1.02e8), n))
    lek = np.log(ek)
    u = rng.uniform(size=n)
    out.update(lk_imc=imc, lk_ekin=ek, lk_lekin=lek, lk_u=u)
    for isel, tag in ((True, "em"), (False, "ep")):
        out[f"lk_{tag}"] = R.electron_lookups(imc, ek, lek, isel)
        out[f"sx_{tag}"] = R.electron_stepping_xsecs(imc, ek, lek, isel)
    mx, pid = R.gamma_lookups(imc, ek, lek, u)
    out.update(gm_mxsec=mx, gm_pid=pid)
    couples = rng.choice(np.array([5, 6], dtype=np.int32), n).astype(np.int32)
    mats = rng.choice(np.array([3, 4], dtype=np.int32), n).astype(np.int32)
    out.update(sel_couples=couples, sel_mats=mats)
    for kind, idx in ((0, couples), (1, couples), (2, mats)):
        for isel, tag in ((True, "em"), (False, "ep")):
            out[f"sel_{kind}_{tag}"] = R.select_target_element(kind, isel, idx, ek, lek, u)
    np.savez_compressed(os.path.join(HERE, "lookups.npz"), **out)

    n = 2048
    out = {}
    b = batches.make_electron_batch(n, ft.num_matcut, seed=909)
    out.update(batch_arrays(b, "in_"))
    h = b.copy()
    R.electron_howfar(h, SEED, 1)
    out.update(batch_arrays(h, "howfar_"))
    # geometry stub between the two calls: every 5th step is halved and ends on a boundary
    cut = (np.arange(n) % 5) == 0
    h.gstep_pstep[cut, 0] *= 0.5
    h.meta[:, 1] = np.where(cut, h.meta[:, 1] | 0x02, h.meta[:, 1] & ~0x02)
    out.update(batch_arrays(h, "geom_"))
    q = batches.SecondaryHostQueue(2 * n)
    R.electron_perform(h, q, SEED, 1)
    out.update(batch_arrays(h, "perform_"))
    out.update(sec_arrays(q, "perform_sec_"))
    s = b.copy()
    q = batches.SecondaryHostQueue(2 * n)
    R.electron_step(s, q, SEED, 1)
    out.update(batch_arrays(s, "step_"))
    out.update(sec_arrays(q, "step_sec_"))
    np.savez_compressed(os.path.join(HERE, "electron_steps.npz"), **out)

    out = {}
    g = batches.make_gamma_batch(n, ft.num_matcut, seed=910, boundary_fraction=0.1)
    out.update(batch_arrays(g, "in_"))
    s = g.copy()
    q = batches.SecondaryHostQueue(2 * n)
    R.gamma_step(s, q, SEED, 1)
    out.update(batch_arrays(s, "step_"))
    out.update(sec_arrays(q, "step_sec_"))
    np.savez_compressed(os.path.join(HERE, "gamma_steps.npz"), **out)
    for f in ("lookups.npz", "electron_steps.npz", "gamma_steps.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
