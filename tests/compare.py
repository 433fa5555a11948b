"""Batch comparison helpers shared by the parity tests.

Tolerances (BASELINE.json north_star): indices / counts / flags bit exact; energies, step lengths and
directions within 1e-12 relative in FP64 (direction components: 1e-12 absolute, they can be ~0).
"""
import numpy as np

REL = 1.0e-12
ABS_DIR = 1.0e-12


def rel_close(a, b, rel=REL, abs_tol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    diff = np.abs(a - b)
    with np.errstate(invalid="ignore"):
        ok = (diff <= rel * np.maximum(np.abs(a), np.abs(b)) + abs_tol) | (a == b)
    # a NaN never matches (not even another NaN): compared state must be finite
    return ok & ~(np.isnan(a) | np.isnan(b))


def compare_group(name, a, b, kinds, report, mask=None):
    """a, b: (n,2) arrays; kinds: per column 'rel' | 'dir' | 'exact' | None"""
    for col, kind in enumerate(kinds):
        if kind is None:
            continue
        x, y = a[:, col], b[:, col]
        if mask is not None:
            x, y = x[mask], y[mask]
        if kind == "exact":
            ok = (x == y) | (np.isnan(x) & np.isnan(y))
        elif kind == "dir":
            ok = rel_close(x, y, REL, ABS_DIR)
        else:
            ok = rel_close(x, y)
        nbad = int((~ok).sum())
        nexact = int(((x == y) | (np.isnan(x) & np.isnan(y))).sum())
        report[f"{name}[{col}]"] = dict(bad=nbad, n=len(x), bit_exact=nexact,
                                        first_bad=(int(np.flatnonzero(~ok)[0]) if nbad else -1))


ELECTRON_KINDS = dict(
    ekin_logekin=("rel", "rel"), dirx_diry=("dir", "dir"), dirz_safety=("dir", "exact"), nia01=("rel", "rel"),
    nia23=("rel", "rel"), msc_irange_dynrf=("rel", "rel"), msc_tlimmin_gauss=("rel", None),
    gstep_pstep=("rel", "rel"), edep_dispx=("rel", None), dispy_dispz=(None, None),
    mfp01=("rel", "rel"), mfp23=("rel", "rel"), range_lambtr1=("rel", "rel"), tstep_zpath=("rel", "rel"),
    par12=("rel", "rel"), par3_pad=("rel", None),
)
GAMMA_KINDS = dict(
    ekin_logekin=("rel", "rel"), dirx_diry=("dir", "dir"), dirz_nia0=("dir", "rel"), gstep_mfp0=("rel", "rel"),
    edep_pemxsec=("rel", None),
)


def compare_electron_batches(a, b, handover=True):
    rep = {}
    for g, kinds in ELECTRON_KINDS.items():
        if not handover and g in ("mfp01", "mfp23", "range_lambtr1", "tstep_zpath", "par12", "par3_pad"):
            continue
        compare_group(g, getattr(a, g), getattr(b, g), kinds, rep)
    # the MSC displacement is a rotated vector (fDisplacement, G4HepEmElectronManager.icc:300-321): its components
    # cancel, so the tolerance is relative to the length of the vector (plus the 1e-12 absolute of directions)
    da = np.stack([a.edep_dispx[:, 1], a.dispy_dispz[:, 0], a.dispy_dispz[:, 1]], axis=1)
    db = np.stack([b.edep_dispx[:, 1], b.dispy_dispz[:, 0], b.dispy_dispz[:, 1]], axis=1)
    norm = np.maximum(np.linalg.norm(da, axis=1), np.linalg.norm(db, axis=1))
    both_nan = np.isnan(da) & np.isnan(db)
    ok = ((np.abs(da - db) <= (REL * norm + ABS_DIR)[:, None]) | (da == db) | both_nan).all(axis=1)
    rep["displacement"] = dict(bad=int((~ok).sum()), n=a.n, bit_exact=int(((da == db) | both_nan).all(axis=1).sum()),
                               first_bad=(int(np.flatnonzero(~ok)[0]) if (~ok).any() else -1))
    # the cached Gaussian variate only matters while its flag is set
    fa = a.meta[:, 1]
    cached = (fa & 0x40) != 0
    compare_group("gauss", a.msc_tlimmin_gauss, b.msc_tlimmin_gauss, (None, "rel"), rep, mask=cached)
    for col, nm in enumerate(("imc", "flags", "id", "draws")):
        bad = int((a.meta[:, col] != b.meta[:, col]).sum())
        rep[f"meta.{nm}"] = dict(bad=bad, n=a.n, bit_exact=a.n - bad,
                                 first_bad=(int(np.flatnonzero(a.meta[:, col] != b.meta[:, col])[0]) if bad else -1))
    bad = int((a.winner != b.winner).sum())
    rep["winner"] = dict(bad=bad, n=a.n, bit_exact=a.n - bad, first_bad=(int(np.flatnonzero(a.winner != b.winner)[0]) if bad else -1))
    return rep


def compare_gamma_batches(a, b, pe_mask=None):
    rep = {}
    for g, kinds in GAMMA_KINDS.items():
        compare_group(g, getattr(a, g), getattr(b, g), kinds, rep)
    # fPEmxSec is defined when PE was selected or the photon was below the upper edge of the second energy window
    # (gm_emax1 = 2 m_e c^2; G4HepEmGammaManager.icc:118-139,182): callers that know the pre-step energies pass the mask
    pe_defined = (a.winner == 2) if pe_mask is None else ((a.winner == 2) | pe_mask)
    compare_group("pemxsec", a.edep_pemxsec, b.edep_pemxsec, (None, "rel"), rep, mask=pe_defined)
    for col, nm in enumerate(("imc", "flags", "id", "draws")):
        bad = int((a.meta[:, col] != b.meta[:, col]).sum())
        rep[f"meta.{nm}"] = dict(bad=bad, n=a.n, bit_exact=a.n - bad,
                                 first_bad=(int(np.flatnonzero(a.meta[:, col] != b.meta[:, col])[0]) if bad else -1))
    bad = int((a.winner != b.winner).sum())
    rep["winner"] = dict(bad=bad, n=a.n, bit_exact=a.n - bad, first_bad=(int(np.flatnonzero(a.winner != b.winner)[0]) if bad else -1))
    return rep


def compare_secondaries(qa, qb):
    ra, rb = qa.sorted_records(), qb.sorted_records()
    rep = {}
    na, nb = len(ra["ekin"]), len(rb["ekin"])
    rep["count"] = dict(bad=int(na != nb), n=1, bit_exact=int(na == nb), first_bad=-1, a=na, b=nb)
    if na != nb:
        return rep
    for k in ("parent_index", "slot", "parent_id", "kind"):
        bad = int((ra[k] != rb[k]).sum())
        rep[k] = dict(bad=bad, n=na, bit_exact=na - bad, first_bad=(int(np.flatnonzero(ra[k] != rb[k])[0]) if bad else -1))
    ok = rel_close(ra["ekin"], rb["ekin"])
    rep["ekin"] = dict(bad=int((~ok).sum()), n=na, bit_exact=int((ra["ekin"] == rb["ekin"]).sum()),
                       first_bad=(int(np.flatnonzero(~ok)[0]) if (~ok).any() else -1))
    ok = rel_close(ra["dir"], rb["dir"], REL, ABS_DIR).all(axis=1)
    rep["dir"] = dict(bad=int((~ok).sum()), n=na, bit_exact=int((ra["dir"] == rb["dir"]).all(axis=1).sum()),
                      first_bad=(int(np.flatnonzero(~ok)[0]) if (~ok).any() else -1))
    return rep


def total_bad(rep):
    return sum(v["bad"] for v in rep.values())


def format_report(rep, only_bad=False):
    lines = []
    for k, v in rep.items():
        if only_bad and v["bad"] == 0:
            continue
        lines.append(f"{k:24s} bad={v['bad']:8d} / {v['n']:8d}  bit-exact={v['bit_exact']:8d}  first_bad={v['first_bad']}")
    return "\n".join(lines)
