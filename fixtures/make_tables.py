#!/usr/bin/env python3
"""Synthetic-but-physically-shaped G4HepEm table set, written in the reference's JSON schema.

The reference builds its tables from Geant4 models + G4EMLOW data (G4HepEmInit); neither is
available offline, and the reference ships no data dump.  Parity of the stepping path needs
*identical inputs*, not physically exact ones, so this script emits a `G4HepEmState` JSON
(`{fParameters, fData}`; schema of G4HepEmDataJsonIO/src/G4HepEmDataJsonIOImpl.hh:127-977)
whose arrays have the layouts the reference's run-time code indexes:

  * e-loss     : G4HepEmInit/src/G4HepEmElectronTableBuilder.cc:40-177  (range|dedx|inv-range SD)
  * lambda     : ...TableBuilder.cc:184-362 (per couple: ioni block, brem block, 5-value headers)
  * e-nuclear  : ...TableBuilder.cc:365-431 (128 pts, 100 MeV - 100 TeV)
  * tr1        : ...TableBuilder.cc:434-498
  * selectors  : ...TableBuilder.cc:501-682, G4HepEmGammaTableBuilder.cc:188-267
  * SB tables  : ...TableBuilder.cc:685-834 (+ G4HepEmSBBremTableBuilder.cc grid conventions)
  * gamma      : G4HepEmGammaTableBuilder.cc:30-184 (3 energy windows, 2/3/9 values per point)
  * material / element derived constants: G4HepEmMaterialInit.cc:116-233

Cross-section *shapes* are textbook formulas (Moller/Bhabha restricted cross sections,
Klein-Nishina, Bethe-Heitler-like logarithmic rise, Highland-like transport mfp, Sandia-like
piecewise a_k/E^k photo-absorption) so that step lengths, branch populations and rejection
rates are calorimeter-like.  Couples: the ATLASbar set (Galactic, G4_Pb, G4_lAr) in two
regions plus PbWO4 (3 elements) and water (2 elements, Z<5 branch) for the element selectors.

Units: MeV, mm (Geant4 internal units).
Usage: python fixtures/make_tables.py [out.json]
"""
import json
import math
import sys

import numpy as np
from scipy.interpolate import CubicSpline

# ---- physical constants (CLHEP values in MeV/mm units) -------------------------------------
MC2 = 0.51099890999999997
ALPHA = 7.2973525653052150e-03
R0 = 2.8179403262e-12  # classical electron radius [mm]
PIR02 = 2.4946724123674787e-23
MIGDAL = 5.2804955733859579e-30
TWOPI_R02_MC2 = 2.0 * PIR02 * MC2
AVOGADRO = 6.02214076e23

N_LOSS_BINS = 84
E_MIN_LOSS = 1.0e-4
E_MAX_LOSS = 1.0e8
BREM_MODEL_LIM = 1000.0
TRACKING_CUT = 1.0e-3


def log_grid(emin, emax, n):
    """FillLogarithmicGrid: returns (grid, log(emin), 1/delta)."""
    lmin = math.log(emin)
    delta = math.log(emax / emin) / (n - 1.0)
    grid = np.exp(lmin + delta * np.arange(n))
    grid[0] = emin
    grid[-1] = emax
    return grid, lmin, 1.0 / delta


def second_derivs(x, y):
    """Second derivatives of the not-a-knot cubic spline through (x, y)."""
    cs = CubicSpline(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), bc_type="not-a-knot")
    return cs(x, 2)


# ---- materials ------------------------------------------------------------------------------
# name, density [g/cm3], [(Z, A, atoms per molecule)], radiation length [mm], mean exc. energy [MeV]
MATERIALS = [
    ("Galactic", 1.0e-25, [(1, 1.00794, 1)], 6.3e27, 21.8e-6),
    ("G4_Pb", 11.35, [(82, 207.217, 1)], 5.6125, 823.0e-6),
    ("G4_lAr", 1.396, [(18, 39.948, 1)], 140.03, 188.0e-6),
    ("G4_PbWO4", 8.28, [(82, 207.217, 1), (74, 183.84, 1), (8, 15.9994, 4)], 8.9034, 600.7e-6),
    ("G4_WATER", 1.0, [(1, 1.00794, 2), (8, 15.9994, 1)], 360.84, 78.0e-6),
]

# couples: (material index, region index, e- cut, e+ cut, gamma cut) [MeV]
COUPLES = [
    (0, 0, 0.00099, 0.00099, 0.00099),
    (1, 0, 1.00386, 0.951321, 0.101843),
    (2, 0, 0.342545, 0.334796, 0.00617835),
    (1, 1, 0.240331, 0.234348, 0.0292362),
    (2, 1, 0.0869021, 0.0861167, 0.00232932),
    (3, 1, 0.877456, 0.838632, 0.0783712),
    (4, 2, 0.351877, 0.342545, 0.00293964),
]

REGIONS = [
    dict(fFinalRange=1.0, fDRoverRange=0.2, fLinELossLimit=0.01, fMSCRangeFactor=0.04, fMSCSafetyFactor=0.6,
         fIsMSCMinimalStepLimit=False, fIsELossFluctuation=True, fIsMultipleStepsInMSCTrans=True, fIsApplyCuts=True),
    dict(fFinalRange=0.1, fDRoverRange=0.15, fLinELossLimit=0.01, fMSCRangeFactor=0.06, fMSCSafetyFactor=0.6,
         fIsMSCMinimalStepLimit=False, fIsELossFluctuation=True, fIsMultipleStepsInMSCTrans=True, fIsApplyCuts=False),
    dict(fFinalRange=1.0, fDRoverRange=0.2, fLinELossLimit=0.02, fMSCRangeFactor=0.2, fMSCSafetyFactor=0.6,
         fIsMSCMinimalStepLimit=True, fIsELossFluctuation=False, fIsMultipleStepsInMSCTrans=False, fIsApplyCuts=True),
]

KSHELL = {1: 13.6e-6, 8: 543.1e-6, 18: 3.2059e-3, 74: 69.525e-3, 82: 88.0045e-3}
FEL_LOW = [0.0, 5.3104, 4.7935, 4.7402, 4.7112]
FINEL_LOW = [0.0, 5.9173, 5.6125, 5.5377, 5.4728]


def coulomb_corr(z):
    az2 = (ALPHA * z) ** 2
    az4 = az2 * az2
    return (1.0 / (1.0 + az2) + 0.20206 - 0.0369 * az2 + 0.0083 * az4 - 0.002 * az2 * az4) * az2


def sandia_per_atom(z):
    """Sandia-like photo-absorption parametrisation per atom: interval lower edges [MeV] and
    4 coefficients a_k (sigma = sum_k a_k / E^k, [mm^2]).  Edges mimic shell structure."""
    k = KSHELL[z]
    edges = [1.0e-5]
    if z >= 8:
        edges.append(max(k / 7.5, 2.0e-5))
    if z >= 18:
        edges.append(k / 3.1)
    if k > 1.0e-5:
        edges.append(k)
    edges.append(max(0.5, 6.0 * k))
    edges = sorted(set(edges))
    coefs = []
    barn = 1.0e-22  # mm^2
    for i, e0 in enumerate(edges):
        # sigma ~ s0 (e0/E)^3 [1 + 0.3 e0/E]; jumps x(4+i) across edges
        jump = 1.0 + 0.9 * i
        s0 = barn * 3.0e1 * (z ** 4.5) * jump * (1.0e-3 / max(e0, 1.0e-3)) ** 0.4 * 1.0e-7
        a3 = s0 * (2.0e-2) ** 3 * 1.0e3
        a1 = 2.0e-4 * a3 / (0.05 + e0) ** 2
        a2 = 0.05 * a3 / (0.02 + e0)
        a4 = 0.3 * a3 * min(e0, 0.05)
        coefs.append([a1, a2, a3, a4])
    return edges, coefs


def build_elements_and_materials():
    elems = {}
    mats = []
    for im, (name, rho, comp, radlen, iexc) in enumerate(MATERIALS):
        mol_a = sum(a * n for (_, a, n) in comp)
        nmol = rho * AVOGADRO / mol_a * 1.0e-3  # per mm^3
        zs = [z for (z, _, _) in comp]
        nat = [nmol * n for (_, _, n) in comp]
        nel = sum(z * n for z, n in zip(zs, nat))
        zeff = sum(z * (a * n) / mol_a for (z, a, n) in comp)
        zeff16 = zeff ** (1.0 / 6.0)
        zeff13 = zeff16 * zeff16
        zsqrt = math.sqrt(zeff)
        dum0 = 9.90395e-1 + zeff16 * (-1.68386e-1 + zeff16 * 9.3286e-2)
        # material level Sandia table: union of the element edges, coefficients = sum n_i a_k^(i)
        per_atom = {z: sandia_per_atom(z) for z in zs}
        edges = sorted(set(e for z in zs for e in per_atom[z][0]))
        mcoefs = []
        for e0 in edges:
            c = np.zeros(4)
            for z, n in zip(zs, nat):
                ez, cz = per_atom[z]
                j = max(i for i, ee in enumerate(ez) if ee <= e0 + 1e-15)
                c += n * np.asarray(cz[j])
            mcoefs.extend(c.tolist())
        mats.append(dict(
            name=name,
            fG4MatIndex=im, fElementVect=zs, fNumOfAtomsPerVolumeVect=nat,
            fDensity=rho * 6.241509074e15, fDensityCorfactor=MIGDAL * nel, fElectronDensity=nel,
            fRadiationLength=radlen, fMeanExEnergy=iexc,
            fSandiaEnergies=edges, fSandiaCoefficients=mcoefs,
            fZeff=zeff, fZeff23=zeff13 * zeff13, fZeffSqrt=zsqrt,
            fUMSCPar=9.62800e-1 - 8.4848e-2 * zsqrt + 4.3769e-3 * zeff,
            fUMSCStepMinPars=[2.7725e1 / (1.0 + 2.03e-1 * zeff), 6.152 / (1.0 + 1.11e-1 * zeff)],
            fUMSCTailCoeff=[2.3785 - zeff13 * (4.1981e-1 - zeff13 * 6.3100e-2),
                            4.7526e-1 + zeff13 * (1.7694 - zeff13 * 3.3885e-1),
                            2.3683e-1 - zeff13 * (1.8111 - zeff13 * 3.2774e-1),
                            1.7888e-2 + zeff13 * (1.9659e-2 - zeff13 * 2.6664e-3)],
            fUMSCThetaCoeff=[dum0 * (1.0 - 8.7780e-2 / zeff), dum0 * (4.0780e-2 + 1.7315e-4 * zeff)],
        ))
        for z in zs:
            if z in elems:
                continue
            dz = float(z)
            logz = math.log(dz)
            fc = coulomb_corr(dz)
            fel = FEL_LOW[z] if z < 5 else math.log(184.15) - logz / 3.0
            finel = FINEL_LOW[z] if z < 5 else math.log(1194.0) - 2.0 * logz / 3.0
            z23 = dz ** (2.0 / 3.0)
            vars1 = z23 / (184.15 * 184.15)
            ez, cz = per_atom[z]
            elems[z] = dict(
                fZet=dz, fZet13=dz ** (1.0 / 3.0), fZet23=z23, fCoulomb=fc, fLogZ=logz,
                fZFactor1=(fel - fc) + finel / dz,
                fDeltaMaxLow=math.exp((42.038 - 8.0 * logz / 3.0) / 8.29) - 0.958,
                fDeltaMaxHigh=math.exp((42.038 - 8.0 * (logz / 3.0 + fc)) / 8.29) - 0.958,
                fILVarS1=1.0 / math.log(vars1), fILVarS1Cond=1.0 / math.log(math.sqrt(2.0) * vars1),
                fSandiaEnergies=list(ez), fSandiaCoefficients=[c for row in cz for c in row],
                fKShellBindingEnergy=KSHELL[z],
            )
    return elems, mats


# ---- model formulas ---------------------------------------------------------------------------
def xsec_ioni_per_electron(ekin, cut, iselectron):
    """Restricted Moller / Bhabha cross section per electron [mm^2] (delta-ray energy > cut)."""
    tmax = 0.5 * ekin if iselectron else ekin
    if cut >= tmax:
        return 0.0
    xmin = cut / ekin
    xmax = tmax / ekin
    tau = ekin / MC2
    gam = tau + 1.0
    gamma2 = gam * gam
    beta2 = tau * (tau + 2.0) / gamma2
    if iselectron:
        g = (2.0 * gam - 1.0) / gamma2
        cross = ((xmax - xmin) * (1.0 - g + 1.0 / (xmin * xmax) + 1.0 / ((1.0 - xmin) * (1.0 - xmax)))
                 - g * math.log(xmax * (1.0 - xmin) / (xmin * (1.0 - xmax)))) / beta2
    else:
        y = 1.0 / (1.0 + gam)
        y2 = y * y
        y12 = 1.0 - 2.0 * y
        b1 = 2.0 - y2
        b2 = y12 * (3.0 + y2)
        y122 = y12 * y12
        b4 = y122 * y12
        b3 = b4 + y122
        cross = ((xmax - xmin) * (1.0 / (beta2 * xmin * xmax) + b2 - 0.5 * b3 * (xmin + xmax)
                                  + b4 * (xmin * xmin + xmin * xmax + xmax * xmax) / 3.0)
                 - b1 * math.log(xmax / xmin))
    return max(0.0, TWOPI_R02_MC2 * cross / ekin)


def dedx_ioni(ekin, cut, nel, iexc, iselectron):
    """Berger-Seltzer-like restricted collision stopping power [MeV/mm]."""
    tau = ekin / MC2
    gam = tau + 1.0
    beta2 = tau * (tau + 2.0) / (gam * gam)
    tmax = 0.5 * ekin if iselectron else ekin
    d = min(cut, tmax) / MC2
    eexc2 = (iexc / MC2) ** 2
    if iselectron:
        f = (-1.0 - beta2 + math.log((tau - d) * d) + tau / (tau - d)
             + (0.5 * d * d + (2.0 * tau + 1.0) * math.log(1.0 - d / tau)) / (gam * gam))
    else:
        y = 1.0 / (1.0 + gam)
        f = (math.log(tau * d) - beta2 * (tau + 2.0 * d - 1.5 * d * d * y) / tau)
    val = math.log(2.0 * (tau + 2.0) / eexc2) + f
    # density-effect like damping at high energy
    x = math.log10(math.sqrt(tau * (tau + 2.0)))
    if x > 0.2:
        val -= min(4.606 * (x - 0.2) * (1.0 - math.exp(-0.8 * x)), 0.85 * val)
    val = max(val, 0.05 * math.log(2.0 * (tau + 2.0) / eexc2) + 0.5)
    # low energy: go to ~sqrt(E) behaviour (dE/dx ~ beta)
    low = 10.0 * iexc
    if ekin < low:
        taul = low / MC2
        b2l = taul * (taul + 2.0) / ((taul + 1.0) ** 2)
        return dedx_ioni(low, cut, nel, iexc, iselectron) * math.sqrt(ekin / low) * 1.0 + 0.0 * b2l
    return TWOPI_R02_MC2 * nel * val / beta2


def brem_shape(ekin, gcut, zs, nat, positron):
    """(restricted dedx, cross section) of bremsstrahlung [MeV/mm, 1/mm]; logarithmic rise, Z(Z+1) scaling."""
    if ekin <= gcut:
        dedx = sum(n * z * (z + 1.0) for z, n in zip(zs, nat)) * 4.0 * ALPHA * R0 * R0 * ekin * 2.0
        return dedx * (ekin / gcut) ** 0.0, 0.0
    etot = ekin + MC2
    dedx = 0.0
    xs = 0.0
    for z, n in zip(zs, nat):
        lrad = math.log(184.15 / z ** (1.0 / 3.0))
        # screening reduces the log at low energy
        scr = lrad * (1.0 - math.exp(-0.35 * math.sqrt(ekin))) + 0.6 * (1.0 - math.exp(-ekin / 0.05)) + 0.15
        fac = 4.0 * ALPHA * R0 * R0 * z * (z + 1.0) * scr
        kappa = gcut / ekin
        xs_z = fac * ((4.0 / 3.0) * (-math.log(kappa) - (1.0 - kappa)) + 0.5 * (1.0 - kappa * kappa))
        de_z = fac * etot * (min(1.0, kappa) * (4.0 / 3.0 - 2.0 / 3.0 * kappa + 0.5 * kappa * kappa))
        if positron:
            # e+ suppression at low energy
            sup = 1.0 - math.exp(-1.0e1 * ekin / (z * z * 1.0e-3) ** 0.5) * 0.7
            xs_z *= sup
            de_z *= sup
        dedx += n * de_z
        xs += n * xs_z
    # high energy LPM like flattening
    xs *= 1.0 / (1.0 + (ekin / 3.0e7) ** 0.5)
    return dedx, max(xs, 0.0)


def tr1_mxsec(ekin, radlen, zeff, positron):
    """First transport macroscopic cross section [1/mm]: Highland-like 1/lambda_1."""
    pbeta = ekin * (ekin + 2.0 * MC2) / (ekin + MC2)
    es = 15.0
    corr = 1.0 + 0.12 * math.log1p(1.0 / (ekin + 1.0e-3)) / (1.0 + 0.02 * zeff)
    val = 0.5 * (es / pbeta) ** 2 / radlen * corr
    # saturate at low energy where lambda1 would go below atomic distances
    val = val / (1.0 + val * 2.0e-6)
    if positron:
        val *= 1.0 - 0.25 * math.exp(-ekin / (0.02 * zeff ** 0.5))
    return val


def nuc_mxsec(ekin, zs, nat, positron):
    s = 0.0
    for z, n in zip(zs, nat):
        a = 2.2 * z if z > 1 else 1.0
        s += n * a * 1.0e-25 * 0.0006 * (1.0 + 0.35 * math.log(ekin / 100.0)) * (1.02 if positron else 1.0)
    return s


def kn_per_electron(e):
    k = e / MC2
    if k < 1.0e-4:
        return (8.0 / 3.0) * PIR02 * (1.0 - 2.0 * k)
    l = math.log(1.0 + 2.0 * k)
    return 2.0 * PIR02 * ((1.0 + k) / (k * k) * (2.0 * (1.0 + k) / (1.0 + 2.0 * k) - l / k) + l / (2.0 * k)
                          - (1.0 + 3.0 * k) / (1.0 + 2.0 * k) ** 2)


def conv_per_atom(e, z):
    if e <= 2.0 * MC2:
        return 0.0
    x = math.log(e / (2.0 * MC2))
    asym = (7.0 / 9.0) * 4.0 * ALPHA * R0 * R0 * z * (z + 1.0) * math.log(184.15 / z ** (1.0 / 3.0))
    return asym * (1.0 - math.exp(-0.25 * x ** 1.7)) * (1.0 + 0.05 * math.exp(-x) * math.log(z + 1.0))


def sandia_eval(edges, coefs, e):
    j = 0
    if e >= edges[0]:
        for i in range(len(edges) - 1, -1, -1):
            if e >= edges[i]:
                j = i
                break
    c = coefs[4 * j: 4 * j + 4]
    inv = 1.0 / e
    return inv * (c[0] + inv * (c[1] + inv * (c[2] + inv * c[3])))


def gnuc_mxsec(e, zs, nat):
    if e < 10.0:
        return 0.0
    s = 0.0
    for z, n in zip(zs, nat):
        a = 2.2 * z if z > 1 else 1.0
        # giant resonance bump + flat
        s += n * a * 1.0e-25 * (0.12 + 0.9 * math.exp(-((math.log(e / 18.0)) ** 2) / 0.18))
    return s


# ---- electron data -----------------------------------------------------------------------------
def build_electron_data(mats, iselectron):
    ncouple = len(COUPLES)
    nmat = len(mats)
    n = N_LOSS_BINS + 1
    egrid, lmin, ildelta = log_grid(E_MIN_LOSS, E_MAX_LOSS, n)
    scale = math.log(E_MAX_LOSS / E_MIN_LOSS)
    eloss = np.zeros(5 * n * ncouple)
    res_start = []
    res = []
    for imc, (imat, _, elcut, poscut, gcut) in enumerate(COUPLES):
        m = mats[imat]
        zs, nat, nel = m["fElementVect"], m["fNumOfAtomsPerVolumeVect"], m["fElectronDensity"]
        ecut = max(elcut, TRACKING_CUT)
        dedx = np.array([dedx_ioni(e, ecut, nel, m["fMeanExEnergy"], iselectron)
                         + brem_shape(e, gcut, zs, nat, not iselectron)[0] for e in egrid])
        sd_dedx = second_derivs(egrid, dedx)
        cs = CubicSpline(egrid, dedx, bc_type="not-a-knot")
        rng = np.zeros(n)
        rng[0] = 2.0 * egrid[0] / dedx[0]
        glx, glw = np.polynomial.legendre.leggauss(16)
        for i in range(n - 1):
            a, b = egrid[i], egrid[i + 1]
            xi = 0.5 * (b - a) * glx + 0.5 * (b + a)
            rng[i + 1] = rng[i] + 0.5 * (b - a) * float(np.sum(glw / np.maximum(cs(xi), 1e-300)))
        sd_rng = second_derivs(egrid, rng)
        sd_inv = second_derivs(rng, egrid)
        s = 5 * n * imc
        eloss[s + 0: s + 2 * n: 2] = rng
        eloss[s + 1: s + 2 * n: 2] = sd_rng
        eloss[s + 2 * n: s + 4 * n: 2] = dedx
        eloss[s + 2 * n + 1: s + 4 * n: 2] = sd_dedx
        eloss[s + 4 * n: s + 5 * n] = sd_inv
        # restricted macroscopic cross sections: ioni block then brem block
        res_start.append(len(res))
        emin_ioni = 2.0 * ecut if iselectron else ecut
        for emin, which in ((emin_ioni, 0), (gcut, 1)):
            npts = max(4, int(round(n * math.log(E_MAX_LOSS / emin) / scale)) + 1)
            g, l0, ild = log_grid(emin, E_MAX_LOSS, npts)
            if which == 0:
                xs = np.array([nel * xsec_ioni_per_electron(e, ecut, iselectron) for e in g])
            else:
                xs = np.array([brem_shape(e, gcut, zs, nat, not iselectron)[1] for e in g])
            imax = int(np.argmax(xs))
            sd = second_derivs(g, xs)
            res.extend([float(npts), float(g[imax]), float(xs[imax]), l0, ild])
            for e, x, d in zip(g, xs, sd):
                res.extend([float(e), float(x), float(d)])
    # electron-nuclear
    nuc_grid, nuc_lmin, nuc_ild = log_grid(100.0, 1.0e8, 128)
    nuc = []
    tr1 = []
    for m in mats:
        zs, nat = m["fElementVect"], m["fNumOfAtomsPerVolumeVect"]
        y = np.array([nuc_mxsec(e, zs, nat, not iselectron) for e in nuc_grid])
        sd = second_derivs(nuc_grid, y)
        for a, b in zip(y, sd):
            nuc.extend([float(a), float(b)])
        y = np.array([tr1_mxsec(e, m["fRadiationLength"], m["fZeff"], not iselectron) for e in egrid])
        sd = second_derivs(egrid, y)
        for a, b in zip(y, sd):
            tr1.extend([float(a), float(b)])

    # element selectors (only multi-element materials)
    def selector(emin, emax, m, kind, cut):
        zs, nat = m["fElementVect"], m["fNumOfAtomsPerVolumeVect"]
        nb = int(7 * math.log(emax / emin) / (6.0 * math.log(10.0)))
        nb = max(nb, 3) + 1
        g, l0, ild = log_grid(emin, emax, nb)
        out = [float(nb), float(len(zs)), l0, ild]
        for e in g:
            out.append(float(e))
            part = []
            ssum = 0.0
            for z, na in zip(zs, nat):
                if kind == "ioni":
                    x = z * xsec_ioni_per_electron(e, cut, iselectron) * (1.0 + 0.02 * math.log(z) * math.exp(-e))
                else:
                    x = brem_shape(e, cut, [z], [1.0], not iselectron)[1]
                    if kind == "rb":
                        x *= 1.0 + 0.03 * math.log(z) / (1.0 + math.log(e / 1000.0))
                ssum += na * max(x, 0.0)
                part.append(ssum)
            part = part[:-1]
            out.extend([p / ssum if ssum > 0 else p for p in part])
        return out

    sel = {k: ([], []) for k in ("ioni", "sb", "rb")}
    for imc, (imat, _, elcut, poscut, gcut) in enumerate(COUPLES):
        m = mats[imat]
        ecut = max(elcut, TRACKING_CUT)
        if len(m["fElementVect"]) < 2:
            for k in sel:
                sel[k][0].append(-1)
            continue
        specs = (("ioni", 2.0 * ecut if iselectron else ecut, E_MAX_LOSS, ecut),
                 ("sb", gcut, BREM_MODEL_LIM, gcut),
                 ("rb", max(gcut, BREM_MODEL_LIM), E_MAX_LOSS, gcut))
        for k, emin, emax, cut in specs:
            if emin >= emax:
                sel[k][0].append(-1)
            else:
                sel[k][0].append(len(sel[k][1]))
                sel[k][1].extend(selector(emin, emax, m, k, cut))

    def arr(a):
        return list(map(float, a)) if len(a) else None

    return dict(
        fNumMatCuts=ncouple, fNumMaterials=nmat, fELossLogMinEkin=lmin, fELossEILDelta=ildelta,
        fELossEnergyGrid=arr(egrid), fELossData=arr(eloss),
        fResMacXSecStartIndexPerMatCut=res_start, fResMacXSecData=arr(res),
        fENucLogMinEkin=nuc_lmin, fENucEILDelta=nuc_ild, fENucEnergyGrid=arr(nuc_grid), fENucMacXsecData=arr(nuc),
        fTr1MacXSecData=arr(tr1),
        fElemSelectorIoniStartIndexPerMatCut=sel["ioni"][0], fElemSelectorIoniData=arr(sel["ioni"][1]),
        fElemSelectorBremSBStartIndexPerMatCut=sel["sb"][0], fElemSelectorBremSBData=arr(sel["sb"][1]),
        fElemSelectorBremRBStartIndexPerMatCut=sel["rb"][0], fElemSelectorBremRBData=arr(sel["rb"][1]),
    )


# ---- Seltzer-Berger sampling tables ---------------------------------------------------------------
def build_sb_tables(mats, elems):
    n_e, n_k = 65, 54
    el_e = 1.0e-4 * 10.0 ** (np.arange(n_e) / 8.0)
    el_e[0] = 1.0e-4
    lel_e = np.log(el_e)
    log_min = math.log(1.0e-4)
    ildelta = 1.0 / (math.log(1.0e4 / 1.0e-4) / (n_e - 1.0))
    # kappa grid: 1e-12, then log-spaced up to 0.5, then linear-ish approach to 1
    kap = np.concatenate(([1.0e-12], np.exp(np.linspace(math.log(1.0e-7), math.log(0.5), 38)),
                          1.0 - np.exp(np.linspace(math.log(0.4), math.log(1.0e-3), 14)), [1.0]))
    assert len(kap) == n_k and np.all(np.diff(kap) > 0)
    lkap = np.log(kap)
    ncouple = len(COUPLES)
    # gamma cuts per Z (sorted unique) and the couple -> cut index map
    cuts_per_z = {}
    for imc, (imat, _, _, _, gcut) in enumerate(COUPLES):
        for z in mats[imat]["fElementVect"]:
            cuts_per_z.setdefault(z, set()).add(gcut)
    cuts_per_z = {z: sorted(v) for z, v in cuts_per_z.items()}
    start_per_z = [0] * 121
    data = []
    for z in sorted(elems):
        cuts = cuts_per_z[z]
        el_emin = max(TRACKING_CUT, min(cuts))
        imin = int(np.searchsorted(el_e, el_emin, side="left")) - 1
        imax = int(np.searchsorted(el_e, BREM_MODEL_LIM, side="left"))
        start_per_z[z] = len(data)
        ndata = (imax - imin + 1) * (len(cuts) + 3 * n_k) + 4
        data.extend([float(ndata), float(imin), float(imax), float(len(cuts))])
        for ie in range(imin, imax + 1):
            e = el_e[ie]
            # pdf in ln(kappa): p(u) = F(kappa); SB-like shape: mild fall towards the tip, Z/E dependent
            tip = 0.08 + 0.5 * math.exp(-e / (0.02 * z))
            pdf = (1.0 - 0.75 * kap + 0.6 * kap * kap) * (1.0 - np.exp(-(1.0 - kap + 1e-3) / tip)) + 1.0e-3
            pdf *= 1.0 + 0.15 * np.tanh(np.log10(kap + 1e-12) / 4.0 + 1.0)
            # start the cumulative at kappa[1]: essentially no probability below 1e-7
            dx = np.diff(lkap)
            cum = np.zeros(n_k)
            inc = 0.5 * (pdf[1:] + pdf[:-1]) * dx
            inc[0] = 1.0e-9 * inc[1]
            cum[1:] = np.cumsum(inc)
            norm = cum[-1]
            cum /= norm
            pdfn = pdf / norm
            par_a = np.zeros(n_k)
            par_b = np.zeros(n_k)
            for i in range(n_k - 1):
                dc = cum[i + 1] - cum[i]
                r = dc / dx[i]
                b = 1.0 - r * r / (pdfn[i] * pdfn[i + 1])
                a = r / pdfn[i] - b - 1.0
                # keep the rational interpolant monotone
                if not (1.0 + a + b > 0.05) or b < -20 or a < -0.99:
                    a, b = 0.0, 0.0
                par_a[i], par_b[i] = a, b
            # cumulative value at each gamma-cut kappa (1 if production impossible)
            for gc in cuts:
                if e > gc:
                    ck = max(1.0e-12, gc / e)
                    il = int(np.searchsorted(kap, ck, side="left")) - 1 if ck > 1.0e-12 else 0
                    il = min(max(il, 0), n_k - 2)
                    a_, b_ = par_a[il], par_b[il]
                    al = math.log(ck / kap[il]) / math.log(kap[il + 1] / kap[il])
                    val = cum[il]
                    if al != 0.0:
                        dum = a_ * (al - 1.0) - 1.0 - b_
                        if abs(b_) > 1e-14:
                            t = -(dum + math.sqrt(max(dum * dum - 4.0 * b_ * al * al, 0.0))) / (2.0 * b_ * al)
                        else:
                            t = al / (1.0 + a_ * (1.0 - al))
                        t = min(max(t, 0.0), 1.0)
                        val = t * (cum[il + 1] - cum[il]) + cum[il]
                    data.append(float(val))
                else:
                    data.append(1.0)
            for ik in range(n_k):
                data.extend([float(cum[ik]), float(par_a[ik]), float(par_b[ik])])
    gc_start = []
    gc_idx = []
    for imc, (imat, _, _, _, gcut) in enumerate(COUPLES):
        gc_start.append(len(gc_idx))
        for z in mats[imat]["fElementVect"]:
            gc_idx.append(cuts_per_z[z].index(gcut))
    return dict(
        fLogMinElEnergy=log_min, fILDeltaElEnergy=ildelta,
        fElEnergyVect=list(map(float, el_e)), fLElEnergyVect=list(map(float, lel_e)),
        fKappaVect=list(map(float, kap)), fLKappaVect=list(map(float, lkap)),
        fGammaCutIndxStartIndexPerMC=gc_start, fGammaCutIndices=gc_idx,
        fSBStartTablesStartPerZ=start_per_z, fSBTableData=data,
    )


# ---- gamma data -------------------------------------------------------------------------------------
def build_gamma_data(mats):
    g0, l0, d0 = log_grid(1.0e-4, 0.15, 32)
    g1, l1, d1 = log_grid(0.15, 2.0 * MC2, 32)
    g2, l2, d2 = log_grid(2.0 * MC2, 1.0e8, 256)
    per_mat = 2 * 32 + 3 * 32 + 9 * 256
    data = []
    for m in mats:
        zs, nat, nel = m["fElementVect"], m["fNumOfAtomsPerVolumeVect"], m["fElectronDensity"]
        pe = lambda e: max(0.0, sandia_eval(m["fSandiaEnergies"], m["fSandiaCoefficients"], e))
        comp = lambda e: nel * kn_per_electron(e)
        for e in g0:
            data.extend([float(e), comp(e)])
        for e in g1:
            data.extend([float(e), comp(e) + pe(e), pe(e)])
        conv = np.array([sum(n * conv_per_atom(e, z) for z, n in zip(zs, nat)) for e in g2])
        cmp_ = np.array([comp(e) for e in g2])
        pe_ = np.array([pe(e) for e in g2])
        tot = conv + cmp_ + pe_ + np.array([gnuc_mxsec(e, zs, nat) for e in g2])
        sds = [second_derivs(g2, y) for y in (tot, conv, cmp_, pe_)]
        for i, e in enumerate(g2):
            data.append(float(e))
            for y, sd in zip((tot, conv, cmp_, pe_), sds):
                data.extend([float(y[i]), float(sd[i])])
    # conversion element selector
    emin, emax = 2.0 * MC2, 1.0e8
    nconv = int(7 * math.log(emax / emin) / (6.0 * math.log(10.0)))
    gc, lc, dc = log_grid(emin, emax, nconv)
    start = []
    sel = []
    for m in mats:
        zs, nat = m["fElementVect"], m["fNumOfAtomsPerVolumeVect"]
        if len(zs) < 2:
            start.append(-1)
            continue
        start.append(len(sel))
        sel.append(float(len(zs)))
        for e in gc:
            ssum = 0.0
            part = []
            for z, n in zip(zs, nat):
                ssum += n * conv_per_atom(e * 1.0000001, z)
                part.append(ssum)
            sel.extend([p / ssum if ssum > 0.0 else p for p in part[:-1]])
    return dict(
        fNumMaterials=len(mats), fDataPerMat=per_mat, fNumData0=64, fNumData1=96,
        fEMin0=1.0e-4, fEMax0=0.15, fLogEMin0=l0, fEILDelta0=d0,
        fEMax1=2.0 * MC2, fLogEMin1=l1, fEILDelta1=d1,
        fEMax2=1.0e8, fLogEMin2=l2, fEILDelta2=d2,
        fMacXsecData=data,
        fElemSelectorConvLogMinEkin=lc, fElemSelectorConvEILDelta=dc,
        fElemSelectorConvStartIndexPerMat=start, fElemSelectorConvEgrid=list(map(float, gc)),
        fElemSelectorConvData=sel if sel else None,
    )


def build_state():
    elems, mats = build_elements_and_materials()
    matcut = dict(
        fNumG4MatCuts=len(COUPLES), fNumMatCutData=len(COUPLES),
        fG4MCIndexToHepEmMCIndex=list(range(len(COUPLES))),
        fMatCutData=[dict(fSecElProdCutE=max(ec, TRACKING_CUT), fSecPosProdCutE=max(pc, TRACKING_CUT),
                          fSecGamProdCutE=gc, fLogSecGamCutE=math.log(gc),
                          fHepEmMatIndex=im, fG4MatCutIndex=i, fG4RegionIndex=ir)
                     for i, (im, ir, ec, pc, gc) in enumerate(COUPLES)],
    )
    matdata = dict(
        fNumG4Material=len(mats), fNumMaterialData=len(mats),
        fG4MatIndexToHepEmMatIndex=list(range(len(mats))),
        fMaterialData=[{k: v for k, v in m.items() if k != "name"} for m in mats],
    )
    params = dict(
        fElectronTrackingCut=TRACKING_CUT, fGammaTrackingCut=0.0,
        fMinLossTableEnergy=E_MIN_LOSS, fMaxLossTableEnergy=E_MAX_LOSS, fNumLossTableBins=N_LOSS_BINS,
        fElectronBremModelLim=BREM_MODEL_LIM, fIsMSCPositronCor=True, fIsMSCDisplacement=True,
        fNumRegions=len(REGIONS), fParametersPerRegion=REGIONS,
    )
    data = dict(
        fTheMatCutData=matcut, fTheMaterialData=matdata,
        fTheElementData=[elems[z] for z in sorted(elems)],
        fTheElectronData=build_electron_data(mats, True),
        fThePositronData=build_electron_data(mats, False),
        fTheSBTableData=build_sb_tables(mats, elems),
        fTheGammaData=build_gamma_data(mats),
    )
    return dict(fParameters=params, fData=data)


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "tests/golden/hepem_state.json"
    state = build_state()
    with open(out, "w") as f:
        json.dump(state, f, separators=(",", ":"))
    print(f"wrote {out}")


if __name__ == "__main__":
    main()
