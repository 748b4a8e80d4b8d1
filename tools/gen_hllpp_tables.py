"""Monte-Carlo generator for the packaged HyperLogLog++ bias-correction tables.

Why this exists: the reference reads `datasketch.hyperloglog_const._thresholds/_bias/_raw_estimate`
(/root/reference/src/hashing.py:78-80).  `datasketch` is not installed in this image and cannot be
fetched (no network), and its empirical tables (Heule et al. 2013, appendix) cannot be restated from
first principles.  The engine therefore takes the tables as *runtime inputs*: when `datasketch` is
importable its tables are used; otherwise the tables written by this script are used (and the oracle's
`datasketch` stub serves the very same arrays, so parity tests compare like with like).

Method (the one the HLL++ paper describes): for each precision p and each of T interpolation
cardinalities n, simulate HLL sketches of n distinct uniformly-hashed items, average the raw estimate
E = alpha_m * m^2 / sum_j 2^-reg_j, and store (mean E, mean E - n).  A register that received c items
holds max of c geometric ranks, P(reg <= r) = (1 - 2^-r)^c, sampled by inverse CDF.

Deterministic: seed 20131 + p.  Output: subgraph_sketching_b200/data/hllpp_tables.npz
Run:  python tools/gen_hllpp_tables.py            (about ten minutes on 8 cores)
"""
import os
import sys
import numpy as np

# HLL++ switch-over thresholds for p = 4..18 (Heule et al. 2013, section 5.2 / datasketch `_thresholds`)
THRESHOLDS = [10, 20, 40, 80, 220, 400, 900, 1800, 3100, 6500, 11500, 20000, 50000, 120000, 350000]


def alpha(p):
    m = 1 << p
    if p == 4:
        return 0.673
    if p == 5:
        return 0.697
    if p == 6:
        return 0.709
    return 0.7213 / (1.0 + 1.079 / m)


def simulate(p, rng, budget=1 << 22):
    m = 1 << p
    T = {4: 80, 5: 160}.get(p, 200)
    trials = int(max(16, min(4000, budget // m)))
    cards = np.unique(np.round(np.linspace(0, 5 * m, T)).astype(np.int64))
    est = np.zeros(len(cards))
    a = alpha(p)
    max_reg = 64 - p + 1
    pvals = np.full(m, 1.0 / m)
    for i, n in enumerate(cards):
        counts = rng.multinomial(int(n), pvals, size=trials)  # [trials, m]
        u = rng.random(counts.shape)
        with np.errstate(divide='ignore', invalid='ignore'):
            tail = -np.expm1(np.log(u) / counts)  # 1 - U^(1/c)
            reg = np.ceil(-np.log2(tail))
        reg = np.where(counts > 0, np.clip(reg, 1, max_reg), 0.0)
        raw = a * m * m / np.sum(np.exp2(-reg), axis=1)
        est[i] = raw.mean()
    return cards, est, est - cards, trials


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'subgraph_sketching_b200', 'data',
                       'hllpp_tables.npz')
    ps = [int(x) for x in sys.argv[1:]] or list(range(4, 19))
    blob = {}
    if os.path.exists(out):
        blob = dict(np.load(out))
    blob['thresholds'] = np.asarray(THRESHOLDS, dtype=np.int64)
    for p in ps:
        rng = np.random.default_rng(20131 + p)
        cards, est, bias, trials = simulate(p, rng)
        blob[f'raw_estimate_p{p}'] = est.astype(np.float64)
        blob[f'bias_p{p}'] = bias.astype(np.float64)
        blob[f'cards_p{p}'] = cards
        print(f'p={p} T={len(cards)} trials={trials} est[0]={est[0]:.4f} est[-1]={est[-1]:.2f} '
              f'bias[0]={bias[0]:.4f} bias[-1]={bias[-1]:.3f} monotone={bool(np.all(np.diff(est) > 0))}', flush=True)
        np.savez_compressed(out, **blob)


if __name__ == '__main__':
    main()
