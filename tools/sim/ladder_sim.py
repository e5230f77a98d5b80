"""Monte-Carlo of the filter's threshold policy (no GPU): sampled hint, then a ladder of tighter levels
raised at checkpoints from live counts of clean 64-column blocks.  Reports emitted groups, captured elements and
certification failures per row for a cfg2-like row population (Gaussian scores, uniform seen items)."""
import numpy as np, math, sys

def run(nI=26744, k=100, seen_avg=144, rows=2048, n_sample=16, z_hint=4.0, z_levels=(3.0, 2.0, 1.0, 0.0), z_cp=3.0, z_fut=1.0,
        cp_stride=None, seed=0, margin=None, twodelta=0.024, ladder=True, interp=0.0):
    rng = np.random.default_rng(seed)
    n_tiles = (nI + 127) // 128
    pad = n_tiles * 128
    stride = (n_tiles - 1) // n_sample
    sample_tiles = np.arange(n_sample) * stride
    if margin is None: margin = max(16, k // 4)
    need = k + margin
    if cp_stride is None: cp_stride = max(8, min(255, -(-n_tiles // 8)))
    emitted_groups, captured, fails, final_rank = [], [], 0, []
    for r in range(rows):
        s = rng.standard_normal(pad).astype(np.float32)
        s[nI:] = -np.inf
        seen_u = rng.poisson(seen_avg)
        seen = rng.choice(nI, size=min(seen_u, nI), replace=False)
        is_seen = np.zeros(pad, bool); is_seen[seen] = True
        N_un = nI - seen_u
        # hint phase: 32-col block maxima over sample tiles, clean blocks only
        blk32 = s.reshape(-1, 32).max(1); dirty32 = is_seen.reshape(-1, 32).any(1)
        sb = np.concatenate([np.arange(t * 4, t * 4 + 4) for t in sample_tiles])
        vals = np.sort(blk32[sb][~dirty32[sb]])[::-1]
        n = len(vals)
        f0 = need / N_un
        pg = 1.0 - math.exp(32.0 * math.log1p(-f0))
        def rank(z): return int(math.ceil(n * pg + z * math.sqrt(n * pg * (1 - pg)))) + 1
        r0 = rank(z_hint)
        if r0 > n - 4: thr = -np.inf
        else: thr = vals[r0 - 1]
        levels = []
        for z in z_levels:
            rj = max(1, rank(z))
            levels.append(vals[min(rj, n) - 1])
        # main phase
        blk64 = s.reshape(-1, 64).max(1); dirty64 = is_seen.reshape(-1, 64).any(1)
        grp = s.reshape(-1, 8).max(1)
        emitted = 0
        t0 = 0
        cum_clean = 0; cum = np.zeros(len(levels))
        while t0 < n_tiles:
            t1 = min(n_tiles, t0 + cp_stride)
            emitted += int((grp[t0 * 16:t1 * 16] >= thr).sum())
            b = slice(t0 * 2, t1 * 2)
            cl = ~dirty64[b]
            cum_clean += int(cl.sum())
            for j, L in enumerate(levels): cum[j] += int((blk64[b][cl] >= L).sum())
            t0 = t1
            if ladder and t0 < n_tiles and cum_clean >= 16:
                f = t0 / n_tiles
                req = need + z_fut * math.sqrt(need * (1 - f))
                Nlos = []
                for j in range(len(levels)):
                    q = cum[j] / cum_clean
                    se = math.sqrt(max(q * (1 - q), 1e-9) / cum_clean)
                    qlo = q - z_cp * se
                    Nlos.append(-math.log1p(-min(max(qlo, 1e-6), 0.999)) / 64.0 * N_un)
                for j in range(len(levels) - 1, -1, -1):
                    if Nlos[j] >= req:
                        cand = levels[j]
                        if interp and j + 1 < len(levels) and Nlos[j + 1] < req and Nlos[j + 1] > 0:
                            w = (math.log(Nlos[j]) - math.log(req)) / (math.log(Nlos[j]) - math.log(Nlos[j + 1]))
                            cand = levels[j] + (levels[j + 1] - levels[j]) * min(max(w, 0.0), 1.0) * interp
                        if cand > thr: thr = cand
                        break
        un = s[:nI][~is_seen[:nI]]
        cap = int((s[:nI] >= thr).sum())
        un_sorted = np.sort(un)[::-1]
        T = un_sorted[k - 1]
        ok = (T - twodelta) >= thr
        fails += 0 if ok else 1
        emitted_groups.append(emitted); captured.append(cap)
        final_rank.append(int((un >= thr).sum()))
    eg, cp = np.array(emitted_groups), np.array(captured)
    return dict(groups_mean=eg.mean(), groups_p99=np.percentile(eg, 99), groups_max=eg.max(), cap_mean=cp.mean(), cap_p99=np.percentile(cp, 99),
                fail_rate=fails / rows, unmasked_above_mean=np.mean(final_rank))

if __name__ == "__main__":
    print("baseline (no ladder):", run(ladder=False, rows=1024))
    for zc in (2.5, 3.0, 3.5):
        for zf in (1.0, 2.0):
            print("ladder z_cp", zc, "z_fut", zf, run(z_cp=zc, z_fut=zf, rows=1024))
