"""CPU model (numpy, float32 arithmetic) of the p4 forward's per-row online softmax: lazily updated reference max,
speculative exponentials with the old reference, tile-SUM vote, 2-in-8 polynomial exp2 with the exponent add done on the
raw bits (uint32 wrap-around modelled), rescale path.  Used to hunt for range holes of the kind the B200 run found
(profiles/r01s2_rescale_bug.log) without a GPU:  python scripts/model_p4_softmax.py [--no-clamp]

It checks, for adversarial score sequences (huge jumps up and down, -inf masks, ties, tiny and huge magnitudes), that every
released P is finite, non-negative and <= 2^9 (fp16-safe), and that LSE matches a float64 evaluation."""
import sys
import numpy as np

f32 = np.float32
CLAMP = "--no-clamp" not in sys.argv
C2 = f32(1.4426950408889634 / np.sqrt(128.0))
INV_C2 = f32(1.0) / C2
MAGIC = f32(12582912.0)
COEF = [f32(0.05517115816473961), f32(0.2426101416349411), f32(0.6932609677314758), f32(0.9999281167984009)]


def ex2_mufu(x):
    with np.errstate(over="ignore", under="ignore"):
        y = np.exp2(x.astype(np.float64)).astype(f32)
    y[np.abs(y) < f32(1.17549435e-38)] = 0          # ftz
    return y


def ex2_poly(s, neg):
    """the kernel's polynomial path on raw scores s (float32 array) with offset neg"""
    s = s.copy()
    if CLAMP:
        s_floor, s_ceil = (f32(-125.0) - neg) * INV_C2, (f32(126.0) - neg) * INV_C2
        s = np.minimum(np.maximum(s, s_floor), s_ceil)        # fmaxf / fminf drop NaN like the hardware
    else:
        s = np.maximum(s, (f32(-125.0) - neg) * INV_C2)
    with np.errstate(all="ignore"):
        x = (s * C2 + neg).astype(f32)                        # (fma in the kernel; one rounding more here is harmless)
        tt = (x + MAGIC).astype(f32)
        nnf = (tt * f32(-1.0) + MAGIC).astype(f32)
        f = (x + nnf).astype(f32)
        pl = (COEF[0] * f + COEF[1]).astype(f32)
        pl = (pl * f + COEF[2]).astype(f32)
        pl = (pl * f + COEF[3]).astype(f32)
    bits = (pl.view(np.uint32).astype(np.uint64) + ((tt.view(np.uint32).astype(np.uint64) << 23) & 0xFFFFFFFF)) & 0xFFFFFFFF
    return bits.astype(np.uint32).view(f32)


EMU = np.zeros(128, dtype=bool)          # 2 of every 8 column pairs inside each 16-column chunk: pairs 0 and 4
for c in range(128):
    pair_in_chunk = (c % 16) // 2
    EMU[c] = pair_in_chunk in (0, 4)


def exps(s, neg):
    with np.errstate(all="ignore"):
        x = (s * C2 + neg).astype(f32)
    p = ex2_mufu(x)
    p[EMU] = ex2_poly(s[EMU], neg)
    return p


def run_row(tiles):
    """tiles: list of float32[128] raw scores (unscaled q.k), -inf = masked.  Returns (lse, problems)"""
    problems = []
    m_ref, l_a, l_b = f32(-np.inf), f32(0), f32(0)
    for j, s in enumerate(tiles):
        if j == 0:
            m_ref = f32(np.max(s)) if np.any(~np.isnan(s)) else f32(-np.inf)
        neg = f32(0) if m_ref == -np.inf else f32(-m_ref * C2)
        p = exps(s, neg)
        with np.errstate(all="ignore"):
            hs, hp = f32(p[:64].sum(dtype=f32)), f32(p[64:].sum(dtype=f32))
        if j > 0:
            need = (not (hs + hp <= f32(512.0))) or m_ref == -np.inf
            if need:
                mx = f32(np.nanmax(s)) if np.any(~np.isnan(s)) else f32(-np.inf)
                if mx > m_ref:
                    with np.errstate(all="ignore"):
                        alpha = ex2_mufu(np.array([(m_ref - mx) * C2], dtype=f32))[0] if m_ref != -np.inf else f32(0)
                    m_ref = mx
                    l_a, l_b = f32(l_a * alpha), f32(l_b * alpha)
                    neg = f32(-m_ref * C2)
                    p = exps(s, neg)
                    hs, hp = f32(p[:64].sum(dtype=f32)), f32(p[64:].sum(dtype=f32))
        if not np.all(np.isfinite(p)):
            problems.append((j, "non-finite P released"))
        elif p.min() < 0:
            problems.append((j, f"negative P released {p.min()}"))
        elif p.max() > f32(512.0) * f32(1.001):
            problems.append((j, f"P {p.max()} > 2^9 released"))
        l_a, l_b = f32(l_a + hs), f32(l_b + hp)
    l = f32(l_a + l_b)
    scale = f32(1.0 / np.sqrt(128.0))
    lse = f32(0) if (m_ref == -np.inf or not l > 0) else f32(m_ref * scale + np.log(l))
    return lse, problems


def exact_lse(tiles):
    s = np.concatenate(tiles).astype(np.float64) / np.sqrt(128.0)
    s = s[np.isfinite(s)]
    if s.size == 0:
        return 0.0
    m = s.max()
    return m + np.log(np.exp(s - m).sum())


def main():
    rng = np.random.default_rng(0)
    worst, bad, n = 0.0, 0, 0
    for trial in range(4000):
        ntile = int(rng.integers(1, 9))
        kind = trial % 8
        tiles = []
        base = 0.0
        for j in range(ntile):
            sd = float(rng.choice([0.01, 1.0, 30.0, 300.0, 3000.0]))
            if kind == 0:   base += float(rng.choice([0, 50, 500, 5000, 50000]))          # staircase up
            elif kind == 1: base -= float(rng.choice([0, 50, 500, 5000, 50000]))          # staircase down
            elif kind == 2: base = float(rng.normal(0, 20000))                            # random walk of levels
            s = (rng.normal(0, sd, 128) + base).astype(f32)
            if kind == 3:   s[rng.integers(0, 128)] += f32(rng.choice([1500, 15000, 150000]))   # one outlier per tile
            if kind == 4:   s[rng.random(128) < 0.7] = -np.inf                            # heavy masking
            if kind == 5 and j < ntile - 1: s[:] = -np.inf                                # only the last tile is visible
            if kind == 6:   s[:] = f32(base + rng.choice([0.0, 1e-3]))                    # ties
            if kind == 7:   s[64:] = -np.inf                                              # one half of the row masked
            tiles.append(s)
        lse, problems = run_row(tiles)
        ref = exact_lse(tiles)
        err = abs(float(lse) - ref) / max(1.0, abs(ref))
        n += 1
        if problems or err > 2e-4:
            bad += 1
            if bad <= 8:
                print(f"BAD trial {trial} kind {kind}: lse {float(lse):.6g} ref {ref:.6g} rel err {err:.2e} problems {problems[:3]}")
        worst = max(worst, err)
    print(f"model_p4_softmax: clamp={'on' if CLAMP else 'off'}  {n} rows, {bad} bad, worst relative LSE error {worst:.2e}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
