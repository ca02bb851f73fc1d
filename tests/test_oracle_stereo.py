"""CPU pinning of the ComputeStereoMatches oracle (oracle/orb_stereo.cpp): an independent, statement-by-statement Python
restatement of src/Frame.cc:1026-1421 (numpy float32 scalars, Python lists for the row table) must give identical
mvuRight / mvDepth; plus a committed golden fixture."""
import hashlib
import math
import os

import numpy as np

from synth import synth_stereo

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32 = np.float32
POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def _cround(x):  # C round(): half away from zero
    return f32(math.floor(abs(float(x)) + 0.5) * (1.0 if x >= 0 else -1.0))


def _py_stereo(kL, dL, kR, dR, pyrL, pyrR, scale, inv_scale, mbf, mb):
    N, Nr = len(kL), len(kR)
    uRight = np.full(N, -1, np.float32)
    depth = np.full(N, -1, np.float32)
    thOrbDist = (100 + 50) // 2
    nRows = pyrL[0].shape[0]
    rows = [[] for _ in range(nRows)]
    for iR in range(Nr):
        kpY = f32(kR["y"][iR])
        r = f32(f32(2.0) * scale[kR["octave"][iR]])
        maxr, minr = int(math.ceil(f32(kpY + r))), int(math.floor(f32(kpY - r)))
        for yi in range(minr, maxr + 1):
            rows[yi].append(iR)
    mbf, mb = f32(mbf), f32(mb)
    with np.errstate(divide="ignore"):
        maxD = f32(mbf / mb)
    minD = f32(0)
    pairs = []
    for iL in range(N):
        levelL = int(kL["octave"][iL])
        vL, uL = f32(kL["y"][iL]), f32(kL["x"][iL])
        cands = rows[int(vL)]
        if not cands:
            continue
        minU, maxU = f32(uL - maxD), f32(uL - minD)
        if maxU < 0:
            continue
        bestDist, bestIdxR = 100, 0
        for iR in cands:
            o = int(kR["octave"][iR])
            if o < levelL - 1 or o > levelL + 1:
                continue
            uR = f32(kR["x"][iR])
            if uR >= minU and uR <= maxU:
                dist = int(POP[np.bitwise_xor(dL[iL], dR[iR])].sum())
                if dist < bestDist:
                    bestDist, bestIdxR = dist, iR
        if bestDist < thOrbDist:
            uR0 = f32(kR["x"][bestIdxR])
            sf = inv_scale[levelL]
            suL, svL, suR0 = _cround(f32(uL * sf)), _cround(f32(vL * sf)), _cround(f32(uR0 * sf))
            w = L = 5
            imL, imR = pyrL[levelL].astype(np.float32), pyrR[levelL].astype(np.float32)
            cy, cxL, cxR0 = int(svL), int(suL), int(suR0)
            if f32(suR0 + L - w) < 0 or f32(suR0 + L + w + 1) >= imR.shape[1]:
                continue
            IL = imL[cy - w:cy + w + 1, cxL - w:cxL + w + 1]
            IL = IL - IL[w, w]
            bestS, bestinc = 2 ** 31 - 1, 0
            vD = [f32(0)] * (2 * L + 1)
            for inc in range(-L, L + 1):
                IR = imR[cy - w:cy + w + 1, cxR0 + inc - w:cxR0 + inc + w + 1]
                IR = IR - IR[w, w]
                dist = f32(np.abs(IL - IR).sum(dtype=np.float64))
                if dist < bestS:
                    bestS, bestinc = int(dist), inc
                vD[L + inc] = dist
            if bestinc in (-L, L):
                continue
            d1, d2, d3 = vD[L + bestinc - 1], vD[L + bestinc], vD[L + bestinc + 1]
            with np.errstate(divide="ignore", invalid="ignore"):
                deltaR = f32(f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2))))
            if deltaR < -1 or deltaR > 1:
                continue
            bestuR = f32(scale[levelL] * f32(f32(suR0 + f32(bestinc)) + deltaR))
            disparity = f32(uL - bestuR)
            if disparity >= minD and disparity < maxD:
                if disparity <= 0:
                    disparity = f32(0.01)
                    bestuR = f32(float(uL) - 0.01)
                depth[iL] = f32(mbf / disparity)
                uRight[iL] = bestuR
                pairs.append((bestS, iL))
    if not pairs:
        return 0, uRight, depth
    pairs.sort()
    median = f32(pairs[len(pairs) // 2][0])
    thDist = f32(f32(f32(1.5) * f32(1.4)) * median)
    kept = len(pairs)
    for s, i in reversed(pairs):
        if f32(s) < thDist:
            break
        uRight[i] = -1
        depth[i] = -1
        kept -= 1
    return kept, uRight, depth


def _run(oracle, w, h, nfeat, seed, mb):
    left, right = synth_stereo(w, h, seed)
    oL, oR = oracle.extractor(nfeat, 1.2, 8, 20, 7), oracle.extractor(nfeat, 1.2, 8, 20, 7)
    kL, dL = oL(left)
    kR, dR = oR(right)
    n, ur, dp = oracle.compute_stereo_matches(oL, oR, kL, dL, kR, dR, 386.1448, mb)
    return (oL, oR, kL, dL, kR, dR), n, ur, dp


def test_stereo_oracle_vs_python_restatement(oracle):
    for seed, mb in ((4, 0.0), (6, 386.1448 / 30.0)):
        (oL, oR, kL, dL, kR, dR), n, ur, dp = _run(oracle, 480, 320, 600, seed, mb)
        (sc, isc, _, _), _, _ = oL.tables()
        pyrL = [oL.level(l) for l in range(8)]
        pyrR = [oR.level(l) for l in range(8)]
        pn, pur, pdp = _py_stereo(kL, dL, kR, dR, pyrL, pyrR, sc, isc, 386.1448, mb)
        assert pn == n and np.array_equal(pur, ur) and np.array_equal(pdp, dp)
        assert n > 50


def test_stereo_oracle_golden(oracle):
    g = np.load(os.path.join(G, "stereo_640x480_seed9.npz"))
    _, n, ur, dp = _run(oracle, 640, 480, 1000, 9, 0.0)
    assert n == int(g["n"]) and np.array_equal(ur, g["uright"]) and np.array_equal(dp, g["depth"])
