"""Seeded synthetic inputs shared by the golden generators, the parity tests and
bench.py.  Pure numpy; no reference / oracle import here."""
from __future__ import annotations
import numpy as np

N, K, INFO = 1024, 448, 440
SIGMAS = (0.15, 0.3, 0.4, 0.5)


def awgn_llr_set(n: int, seed: int = 7, sigmas=SIGMAS, encode=None):
    """Config 4: random 440-bit payloads, polar-encoded, BPSK (+1 ⇔ bit 1) + AWGN,
    llr = 2·rx/σ² as float32, σ cycling over `sigmas`.  Returns (llr[n,1024] f32, info[n,440] u8)."""
    if encode is None:
        from _polar_host import encode_bits as encode
    rng = np.random.default_rng(seed)
    llr = np.empty((n, N), np.float32)
    info = np.empty((n, INFO), np.uint8)
    for i in range(n):
        sig = sigmas[i % len(sigmas)]
        info[i] = rng.integers(0, 2, INFO, dtype=np.uint8)
        cw = encode(info[i])
        rx = (2.0 * cw.astype(np.float64) - 1.0) + sig * rng.standard_normal(N)
        llr[i] = (2.0 * rx / (sig * sig)).astype(np.float32)
    return llr, info


def detector_like_llr_set(n: int, seed: int = 11):
    """Tie-prone set (SURVEY.md §7 hard part 1): first half AWGN σ=0.7 around random
    codewords, second half `clip(N(0,2), ±12)` float32 (what `_llr` emits on garbage)."""
    from _polar_host import encode_bits
    rng = np.random.default_rng(seed)
    out = np.empty((n, N), np.float32)
    for i in range(n):
        if i < n // 2:
            cw = encode_bits(rng.integers(0, 2, INFO, dtype=np.uint8))
            rx = (2.0 * cw.astype(np.float64) - 1.0) + 0.7 * rng.standard_normal(N)
            out[i] = (2.0 * rx / 0.49).astype(np.float32)
        else:
            out[i] = np.clip(rng.normal(0.0, 2.0, N), -12.0, 12.0).astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------------
# RX / TX clips.  Built with oracle/tx_oracle.py, whose equality
# with the reference embedder is asserted when the golden files are generated.
# ----------------------------------------------------------------------------------------------
FS = 48_000
CLIP_SPECS = {
    # name: (kind, key, seed, secs, start_ctr, slice)   slice = (embed_secs, offset) or None
    "chirp_aa": ("chirp", bytes([0xAA]) * 32, 52, 3.0, 0, None),
    "noise_44": ("noise", bytes([0x44]) * 32, 1, 3.0, 0, None),
    "silence_ee": ("silence", bytes([0xEE]) * 32, 2, 3.0, 0, None),
    "bench_17": ("noise", None, 17, 3.0, None, "bench"),
    "bench_19": ("chirpmix", None, 19, 3.0, None, "bench"),
    "plain_noise": ("noise", bytes([0x11]) * 32, 5, 3.0, None, "nowm"),
    "short_1s": ("noise", bytes([0x22]) * 32, 6, 1.0, 0, None),
}


def bench_key(i: int) -> bytes:
    import hashlib
    return hashlib.sha256(b"echoseal-bench" + int(i).to_bytes(4, "big")).digest()


def host_signal(kind: str, n: int, rng) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / FS
    T = n / FS
    if kind == "chirp":
        return (0.3 * np.cos(2 * np.pi * (300.0 * t + (3500.0 - 300.0) / (2 * T) * t * t))).astype(np.float32)
    if kind == "noise":
        return (0.05 * rng.standard_normal(n)).astype(np.float32)
    if kind == "chirpmix":
        c = 0.3 * np.cos(2 * np.pi * (300.0 * t + (3500.0 - 300.0) / (2 * T) * t * t))
        return (c + 0.02 * rng.standard_normal(n)).astype(np.float32)
    if kind == "silence":
        return np.zeros(n, np.float32)
    raise ValueError(kind)


def make_clip(name: str, embedder_factory=None):
    """Returns (audio float32[n], key32).  embedder_factory(key, seed) -> object with .frame_ctr and
    .process(x); defaults to the TX oracle with seeded randomness."""
    kind, key, seed, secs, start_ctr, mode = CLIP_SPECS[name]
    return make_clip_spec(kind, key, seed, secs, start_ctr, mode, embedder_factory)


def make_clip_spec(kind, key, seed, secs, start_ctr, mode, embedder_factory=None):
    if embedder_factory is None:
        from oracle import tx_oracle as txo
        embedder_factory = lambda k, s: txo.Embedder(k, txo.seeded_rand(s))
    rng = np.random.default_rng(seed)
    n = int(round(secs * FS))
    if mode == "bench":
        key = bench_key(seed)
        r = int(rng.integers(0, 2000))
        n5 = 5 * FS
        host = host_signal(kind, n5, rng)
        tx = embedder_factory(key, seed)
        tx.frame_ctr = r
        wm = tx.process(host)
        off = int(rng.integers(0, n5 - n))
        return np.ascontiguousarray(wm[off:off + n], dtype=np.float32), key
    host = host_signal(kind, n, rng)
    if mode == "nowm":
        return host, key
    tx = embedder_factory(key, seed)
    tx.frame_ctr = start_ctr
    return np.ascontiguousarray(tx.process(host), dtype=np.float32), key


def assert_llr_close(got, ref, what=""):
    """LLR parity (north_star: 1e-4 relative): |got - ref| <= 1e-4 |ref| + 2e-5.  The absolute floor covers values next
    to zero, where (despread - mean) cancels in float32 (rtwm/detector.py:396-405): 2e-5 is 1.7e-6 of the +-12 clip
    range, the size of the float32 rounding the reference itself carries (oracle vs reference: 1e-6)."""
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    lim = 1e-4 * np.abs(ref) + 2e-5
    assert (err <= lim).all(), f"{what}: max err {err.max():.3e} at ref {ref[np.argmax(err - lim)]:.4f}"
    return float(err.max())
