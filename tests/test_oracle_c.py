"""The C restatement used for CPU-baseline timing must agree with the (golden-pinned) Python oracle."""
import numpy as np

from oracle import kvblock_oracle as o
from oracle import kvblock_oracle_c as oc


def test_c_hash_matches_python_oracle(golden):
    t = golden["text"]
    tp = o.TokenProcessor(t["block_size"], t["hash_seed"])
    assert oc.init_hash(tp.init_hash, t["model"]) == tp.get_init_hash(t["model"])
    toks = np.asarray(t["tokens"], dtype=np.uint32)
    keys, _ = oc.hash_batch(toks, np.array([0, toks.size], dtype=np.int64),
                            np.array([tp.get_init_hash(t["model"])], dtype=np.uint64), t["block_size"])
    assert [int(k) for k in keys] == t["request_keys"]
    rng = np.random.default_rng(3)
    for bs in (1, 16, 24, 300):
        prompts = [rng.integers(0, 1 << int(rng.choice([5, 8, 16, 17, 32])), int(rng.integers(0, 5 * bs + 2)),
                                dtype=np.uint64).astype(np.uint32) for _ in range(20)]
        off = np.zeros(21, dtype=np.int64)
        np.cumsum([len(p) for p in prompts], out=off[1:])
        parents = np.array([int(rng.integers(1, 1 << 62)) for _ in prompts], dtype=np.uint64)
        keys, koff = oc.hash_batch(np.concatenate(prompts) if off[-1] else np.zeros(0, np.uint32), off, parents, bs,
                                   threads=2)
        otp = o.TokenProcessor(bs)
        for i, p in enumerate(prompts):
            want = otp.tokens_to_kv_block_keys(int(parents[i]), [int(x) for x in p], "m") or []
            assert [int(k) for k in keys[koff[i]:koff[i + 1]]] == want


def test_c_score_matches_python_oracle():
    lib = oc.load()
    rng = np.random.default_rng(4)
    ix = lib.kvo_index_new(1 << 12)
    oidx = o.InMemoryIndex(pod_cache_size=13)
    chains = [np.array([int(x) for x in rng.integers(1, 1 << 62, 30)], dtype=np.uint64) for _ in range(20)]
    for c in chains:
        for _ in range(3):
            d, pod, tier = int(rng.integers(1, 31)), int(rng.integers(0, 8)), int(rng.integers(0, 2))
            lib.kvo_index_add(ix, c[:d].ctypes.data, d, pod, tier)
            oidx.add(None, [int(k) for k in c[:d]], [o.PodEntry(str(pod), ["gpu", "cpu"][tier])])
    keys = np.concatenate(chains)
    koff = np.arange(0, 30 * 21, 30, dtype=np.int64)
    w = np.ones(256)
    w[1] = 0.8
    out_n = np.zeros(20, np.int32)
    out_p = np.zeros(20 * 13, np.uint16)
    out_s = np.zeros(20 * 13, np.float64)
    lib.kvo_score_batch(ix, keys.ctypes.data, koff.ctypes.data, 20, w.ctypes.data, out_n.ctypes.data, out_p.ctypes.data,
                        out_s.ctypes.data, 2)
    for p, c in enumerate(chains):
        ks = [int(k) for k in c]
        want = o.longest_prefix_score(ks, oidx.lookup(ks))
        got = {str(int(out_p[p * 13 + j])): float(out_s[p * 13 + j]) for j in range(int(out_n[p]))}
        assert got == want
    lib.kvo_index_free(ix)


def test_c_hash_short_parents_and_extras_match_python_oracle():
    """The C restatement is the checker of the GPU fuzz (tests/soak_hash.py) for short parent heads and pre-encoded
    multimodal extras: it must agree with the golden-pinned Python oracle on exactly those."""
    rng = np.random.default_rng(8)
    for bs in (4, 8, 16, 17):
        otp = o.TokenProcessor(bs)
        prompts, parents, feats = [], [], []
        for i in range(24):
            n = int(rng.integers(0, 6 * bs))
            prompts.append(rng.integers(0, 1 << int(rng.choice([5, 8, 16, 17, 32])), n, dtype=np.uint64).astype(np.uint32))
            cands = [0, 1, 23, 24, 255, 256, 65535, 65536, (1 << 32) - 1, 1 << 32, (1 << 64) - 1]
            parents.append(cands[int(rng.integers(0, len(cands)))])
            nk = n // bs
            if i % 3 == 0 or nk == 0:
                feats.append(None)
            else:
                feats.append([None if rng.random() < 0.5 else
                              o.BlockExtraFeatures([o.MMHash("h%d-%d" % (i, int(rng.integers(0, 99)))) for _ in range(int(rng.integers(1, 3)))])
                              for _ in range(nk)])
        off = np.zeros(len(prompts) + 1, dtype=np.int64)
        np.cumsum([len(p) for p in prompts], out=off[1:])
        # pre-encoded X(extra) per block, empty = nil (what the product hands to the device)
        chunks = []
        for p, f in zip(prompts, feats):
            for b in range(len(p) // bs):
                ef = None if f is None else f[b]
                chunks.append(b"" if ef is None else o.encode_extra_suffix(ef))
        eoff = np.zeros(len(chunks) + 1, dtype=np.int64)
        np.cumsum([len(c) for c in chunks], out=eoff[1:])
        extra = np.frombuffer(b"".join(chunks) or b"\0", dtype=np.uint8).copy()
        keys, koff = oc.hash_batch(np.concatenate(prompts) if off[-1] else np.zeros(0, np.uint32), off,
                                   np.asarray(parents, dtype=np.uint64), bs, extra, eoff, threads=3)
        for i, p in enumerate(prompts):
            want = otp.tokens_to_kv_block_keys(parents[i], [int(x) for x in p], "m", feats[i]) or []
            if parents[i] == 0:
                continue       # parent 0 means "start from the model's init hash" in the Python API, not a literal parent
            assert [int(k) for k in keys[koff[i]:koff[i + 1]]] == want, (bs, i)
