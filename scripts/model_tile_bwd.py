"""CPU model of the tile / carry logic of tzk_bwd.cu (tile_update_kernel + carry_combine_kernel): integer "gradients",
so the per-key sums must match a plain group-by exactly and every key must be updated exactly once."""
import numpy as np


def run(keys, g, TP, sentinel):
    n = len(keys)
    n_tiles = (n + TP - 1) // TP
    upd = {}
    cf = [None] * n_tiles
    cl = [None] * n_tiles

    def update(k, acc):
        assert k not in upd, f"key {k} updated twice"
        upd[k] = acc

    for t in range(n_tiles):
        base = t * TP
        cnt = min(TP, n - base)
        sk = [keys[p] if 0 <= p < n else sentinel for p in range(base - 1, base + cnt + 1)]
        k_first, k_last = sk[1], sk[cnt]
        first_cont = t > 0 and sk[0] == k_first
        last_cont = base + cnt < n and sk[cnt + 1] == k_last
        rows = [g[base + i] for i in range(cnt)]
        # level 1 (blocks of 8) and level 2 (blocks of 64), in place like the kernel (reads of a level see only the
        # previous level's values: emulate the barrier with a copy)
        prev = list(rows)
        for i in range(cnt):
            key = sk[i + 1]
            if key == sentinel:
                continue
            head = i == 0 or sk[i] != key
            if not (head or i % 8 == 0):
                continue
            if not (i + 1 < cnt and (i + 1) % 8 != 0 and sk[i + 2] == key):
                continue
            acc = prev[i]
            j = i + 1
            while j < cnt and j % 8 != 0 and sk[j + 1] == key:
                acc += prev[j]
                j += 1
            rows[i] = acc
        prev = list(rows)
        for i in range(cnt):
            key = sk[i + 1]
            if key == sentinel:
                continue
            head = i == 0 or sk[i] != key
            if not (head or i % 64 == 0):
                continue
            j = (i | 7) + 1
            if not (j < cnt and j % 64 != 0 and sk[j + 1] == key):
                continue
            acc = prev[i]
            while j < cnt and j % 64 != 0 and sk[j + 1] == key:
                acc += prev[j]
                j += 8
            rows[i] = acc
        for i in range(cnt):
            key = sk[i + 1]
            if key == sentinel:
                continue
            if i == 0 or sk[i] != key:
                c_l = i == 0 and first_cont
                c_r = key == k_last and last_cont
                kind = 2 if c_r else (1 if c_l else 3)
                acc = rows[i]
                j = (i | 63) + 1
                while j < cnt and sk[j + 1] == key:
                    acc += rows[j]
                    j += 64
                if kind == 1:
                    assert cf[t] is None
                    cf[t] = acc
                elif kind == 2:
                    assert cl[t] is None
                    cl[t] = acc
                else:
                    update(key, acc)
    for t in range(n_tiles - 1):
        pe = (t + 1) * TP
        k0 = keys[pe - 1]
        if keys[pe] != k0 or k0 == sentinel:
            continue
        ps = t * TP
        if t > 0 and keys[ps] == k0 and keys[ps - 1] == k0:
            continue
        lo, step = pe, TP
        hi = lo + step
        while hi < n and keys[hi] == k0:
            lo = hi
            step <<= 1
            hi = lo + step
        hi = min(hi, n)
        while hi - lo > 1:
            mid = (lo + hi) >> 1
            if keys[mid] == k0:
                lo = mid
            else:
                hi = mid
        te = (hi - 1) // TP
        acc = cl[t]
        for tt in range(t + 1, te):
            acc += cl[tt]
        acc += cf[te]
        update(k0, acc)
    return upd


rng = np.random.default_rng(0)
for trial in range(2000):
    TP = int(rng.choice([2, 4, 8, 32, 256]))
    n = int(rng.integers(1, (40 if TP < 256 else 6) * TP))
    nk = int(rng.choice([1, 2, 3, 5, 50, 1000]))
    keys = np.sort(rng.integers(0, nk, size=n))
    n_pad = int(rng.integers(0, 3 * TP)) if rng.random() < 0.5 else 0
    sentinel = 1 << 20
    keys = np.concatenate([keys, np.full(n_pad, sentinel)]).tolist()
    g = rng.integers(-1000, 1000, size=len(keys)).tolist()
    got = run(keys, g, TP, sentinel)
    want = {}
    for k, v in zip(keys, g):
        if k != sentinel:
            want[k] = want.get(k, 0) + v
    assert got == want, (trial, TP, n)
print("tile/carry model OK")
