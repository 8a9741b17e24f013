"""CPU model (numpy, float32 arithmetic) of the forward kernel's per-row online softmax (flash_fwd_p4_sm100.cu, round 2):

  * a row is owned by two threads (key columns [0,64) and [64,128) of every 128-key tile);
  * EXACT step (first step of an item, every step of a retried item): exact row max, exchanged; the reference moves
    only when the max is more than 2^8 above it (lazy rescale);
  * SPECULATIVE step ("late-agreed reference"): exponentials with the reference agreed so far; each thread publishes
    the top 16 bits of its step sum and reads the peer's ONE STEP LATE; if the row sum of step j-1 exceeded 2^9 both
    shift the reference by floor(log2 sum) — identical arithmetic on identical inputs, no vote;
  * overflow inside a single key tile (step sum not <= 2^100 for bf16 P, 2^15 for fp16 P, or inf / NaN) flags the row:
    the kernel then redoes the whole work item with exact steps.  The 2-in-8 polynomial exp2 adds the exponent on the
    raw bits (uint32 wrap-around modelled): without the upper clamp of its argument a 2^141 jump comes out as a tiny
    negative number and the overflow check does NOT trip (the round-1 hole, profiles/r01s2_rescale_bug.log).

    python scripts/model_p4_softmax.py [--no-clamp] [--fp16]

Checked for adversarial score sequences (huge jumps up and down, -inf masks, ties, tiny and huge magnitudes): a row is
either flagged (and its exact redo matches float64) or every released P is finite, non-negative and within the dtype's
range and the LSE matches a float64 evaluation.  Also reports how many rows of N(0, 1)-like data were flagged (must be 0)."""
import sys

import numpy as np

f32 = np.float32
CLAMP = "--no-clamp" not in sys.argv
FP16 = "--fp16" in sys.argv
OVERFLOW_AT = f32(32768.0) if FP16 else f32(2.0 ** 100)
P_MAX = 65504.0 if FP16 else 3.3e38
C2 = f32(1.4426950408889634 / np.sqrt(128.0))
MAGIC = f32(12582912.0)
COEF = [f32(0.05517115816473961), f32(0.2426101416349411), f32(0.6932609677314758), f32(0.9999281167984009)]
LN2 = f32(0.6931471805599453)


def ex2_mufu(x):
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        y = np.exp2(x.astype(np.float64)).astype(f32)
    y[np.abs(y) < f32(1.17549435e-38)] = 0          # ftz
    return y


def ex2_poly(x, spec):
    """the kernel's polynomial path on the exponent argument x (float32 array)"""
    x = np.maximum(x, f32(-125.0))                  # fmaxf drops NaN like the hardware
    if spec and CLAMP:
        x = np.minimum(x, f32(126.0))
    with np.errstate(all="ignore"):
        tt = (x + MAGIC).astype(f32)
        nnf = (tt * f32(-1.0) + MAGIC).astype(f32)
        f = (x + nnf).astype(f32)
        pl = (COEF[0] * f + COEF[1]).astype(f32)
        pl = (pl * f + COEF[2]).astype(f32)
        pl = (pl * f + COEF[3]).astype(f32)
    bits = (pl.view(np.uint32).astype(np.uint64) + ((tt.view(np.uint32).astype(np.uint64) << 23) & 0xFFFFFFFF)) & 0xFFFFFFFF
    return bits.astype(np.uint32).view(f32)


EMU = np.zeros(64, dtype=bool)           # 2 of every 8 column pairs inside each 16-column chunk: pairs 0 and 4
for c in range(64):
    EMU[c] = ((c % 16) // 2) in (0, 4)


def exps(s, neg, spec):
    with np.errstate(all="ignore"):
        x = (s * C2 + neg).astype(f32)
    p = ex2_mufu(x)
    p[EMU] = ex2_poly(x[EMU], spec)
    return p


def top16(x):
    return np.array([x], dtype=f32).view(np.uint32)[0] >> 16


def from16(b):
    return np.array([int(b) << 16], dtype=np.uint32).view(f32)[0]


def run_row(tiles, exact):
    """tiles: list of float32[128] raw scores (unscaled q.k), -inf = masked.  -> (lse, flagged, problems)"""
    problems, flagged = [], False
    neg, has_ref = f32(0), False
    l = [f32(0), f32(0)]
    hs16 = [0, 0]
    for j, s in enumerate(tiles):
        halves = [s[:64], s[64:]]
        if exact or j == 0:
            fin = s[~np.isnan(s)]
            mx = f32(fin.max()) if fin.size else f32(-np.inf)
            with np.errstate(all="ignore"):
                xm = f32(mx * C2 + neg)
            need = (xm > f32(8.0)) if has_ref else (mx > -np.inf)
            if need:
                alpha = ex2_mufu(np.array([-xm], dtype=f32))[0] if has_ref else f32(0)
                neg, has_ref = f32(-mx * C2), True
                l = [f32(l[0] * alpha), f32(l[1] * alpha)]
            spec = False
        else:
            with np.errstate(all="ignore"):
                tot = f32(from16(hs16[0]) + from16(hs16[1]))
            if tot > f32(512.0):
                e = min(int((np.array([tot], dtype=f32).view(np.uint32)[0] >> 23) & 0xFF) - 127, 120)
                sc = np.array([(127 - e) << 23], dtype=np.uint32).view(f32)[0]
                neg = f32(neg - f32(e))
                l = [f32(l[0] * sc), f32(l[1] * sc)]
            spec = True
        for hh in range(2):
            p = exps(halves[hh], neg, spec)
            with np.errstate(all="ignore"):
                hs = f32(p.sum(dtype=f32))
            hs16[hh] = top16(hs)
            if spec and (not (hs <= OVERFLOW_AT) or (not has_ref and hs != 0)):   # no reference yet: only an exact step can establish one
                flagged = True
            elif not flagged:
                if not np.all(np.isfinite(p)):
                    problems.append((j, "non-finite P released by an unflagged row"))
                elif p.min() < 0:
                    problems.append((j, f"negative P released {p.min()}"))
                elif p.max() > P_MAX:
                    problems.append((j, f"P {p.max()} outside the dtype's range"))
            with np.errstate(all="ignore"):
                l[hh] = f32(l[hh] + hs)
    with np.errstate(all="ignore"):
        lt = f32(l[0] + l[1])
        lse = f32(0) if (not has_ref or not lt > 0) else f32(-neg * LN2 + f32(np.log(lt)))
    return lse, flagged, problems


def exact_lse(tiles):
    s = np.concatenate(tiles).astype(np.float64) / np.sqrt(128.0)
    s = s[np.isfinite(s)]
    if s.size == 0:
        return 0.0
    m = s.max()
    return m + np.log(np.exp(s - m).sum())


def main():
    rng = np.random.default_rng(0)
    worst, bad, n, retried = 0.0, 0, 0, 0
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
        lse, flagged, problems = run_row(tiles, exact=False)
        if flagged:                                    # the kernel redoes the item with exact steps
            retried += 1
            lse, _, problems = run_row(tiles, exact=True)
        ref = exact_lse(tiles)
        err = abs(float(lse) - ref) / max(1.0, abs(ref))
        n += 1
        if problems or not err <= 2e-4:
            bad += 1
            if bad <= 8:
                print(f"BAD trial {trial} kind {kind}: lse {float(lse):.6g} ref {ref:.6g} rel err {err:.2e} flagged {flagged} problems {problems[:3]}")
        worst = max(worst, err if err == err else 1.0)
    # well-behaved data (what BASELINE's randn inputs look like after the 1/sqrt(d) scale): never flagged
    benign = 0
    for trial in range(500):
        tiles = [(rng.normal(0, 11.3, 128)).astype(f32) for _ in range(int(rng.integers(2, 33)))]
        lse, flagged, problems = run_row(tiles, exact=False)
        ref = exact_lse(tiles)
        benign += int(flagged)
        if problems or abs(float(lse) - ref) > 2e-4 * max(1.0, abs(ref)):
            bad += 1
    print(f"model_p4_softmax: clamp={'on' if CLAMP else 'off'} dtype={'fp16' if FP16 else 'bf16'}  {n} adversarial rows ({retried} redone exactly), "
          f"{benign} of 500 benign rows flagged, {bad} bad, worst relative LSE error {worst:.2e}")
    return 1 if (bad or benign) else 0


if __name__ == "__main__":
    sys.exit(main())
