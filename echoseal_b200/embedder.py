"""B200-native drop-in for rtwm/embedder.py (WatermarkEmbedder, TxParams): embed-side spreading.

Same class names, constructor arguments, methods and state attributes as the reference
(rtwm/embedder.py:20-168).  Frame synthesis (CRC-8 + polar encode + PN spread + header + zero-state
hop-band band-pass) and the level-controlled mix run in the sm_100a kernels behind the C-ABI; the
payload sealing (ChaCha20-Poly1305), the PN bits (AES-ECB) and the hop band (HMAC) are produced on
the host and fed to the kernels (BASELINE.json north_star).  `EmbedderBank` is the additive batched
form: many concurrent streams per launch.  No CPU fallback."""
from __future__ import annotations
from dataclasses import dataclass, field
import secrets
import numpy as np
import torch

from . import tx_gpu
from .crypto import SecureChannel
from .utils import choose_band_index, db_to_lin, mseq_63

N_DEFAULT = 1024
K_DEFAULT = 448
EPS = 1e-12
MIN_RMS_SILENCE = 1e-4
MIX_HEADROOM = 0.98
HDR_BITS = 16
HDR_REPEAT = 8
HDR_L = 128
FRAME_LEN = 63 + HDR_L + N_DEFAULT


@dataclass(slots=True)
class TxParams:
    """rtwm/embedder.py:20-27"""
    fs: int = 48_000
    target_rel_db: float = -10.0
    floor_rel_dbfs: float = -35.0
    N: int = N_DEFAULT
    K: int = K_DEFAULT
    preamble: np.ndarray = field(default_factory=lambda: mseq_63())


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("echoseal_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class WatermarkEmbedder:
    def __init__(self, key32: bytes, params: TxParams | None = None) -> None:
        self.p = params or TxParams()
        if self.p.N != N_DEFAULT:
            raise ValueError("the B200 path implements Polar(1024, K) only")
        self.sec = SecureChannel(key32)
        self._band_key = getattr(self.sec, "band_key", key32)
        self.frame_ctr = 0
        self._chip_buf: np.ndarray | None = None
        self._session_nonce = secrets.token_bytes(8)
        self._preamble_sy = 2.0 * self.p.preamble.astype(np.float32) - 1.0
        self._hdr_pn_bits = self.sec.pn_bits(0, HDR_L)
        self._hdr_pn_sy = 2.0 * self._hdr_pn_bits.astype(np.float32) - 1.0
        self._hdr_pn_packed = np.packbits(self._hdr_pn_bits)

    # ------------------------------------------------------------------ API
    def process(self, samples: np.ndarray) -> np.ndarray:
        """rtwm/embedder.py:44-75 — one call = one block: block RMS, floor, headroom limiter."""
        if self._chip_buf is None:
            self._chip_buf = np.empty(0, dtype=np.float32)
        x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
        needed = x.size
        if needed == 0:
            return x.copy()
        missing = needed - self._chip_buf.size
        if missing > 0:
            nfr = (missing + FRAME_LEN - 1) // FRAME_LEN
            chips = self._make_frames(nfr)
            self._chip_buf = np.concatenate((self._chip_buf, chips.reshape(-1)))
        chips = self._chip_buf[:needed]
        self._chip_buf = self._chip_buf[needed:]
        dev = _dev()
        xd = torch.from_numpy(x[None]).to(dev)
        cd = torch.from_numpy(np.ascontiguousarray(chips)[None]).to(dev)
        out, _ = tx_gpu.mix(xd, cd, db_to_lin(self.p.target_rel_db), db_to_lin(self.p.floor_rel_dbfs))
        return out[0].cpu().numpy()

    # ------------------------------------------------------------------ internals
    def _frame_inputs(self, ctr: int, payload: bytes):
        band = choose_band_index(self._band_key, ctr)
        pn = self.sec.pn_bytes_batch(np.array([ctr], np.uint64), FRAME_LEN)[0]
        return payload, pn, band, ctr & 0xFFFF

    def _synth(self, items) -> np.ndarray:
        """items: list of (payload55, pn152, band, ctr_lo16) -> float32[F,1215] via K7."""
        dev = _dev()
        tx_gpu.set_filters(self.p.fs, self.p.preamble)
        F = len(items)
        pay = torch.from_numpy(np.frombuffer(b"".join(i[0] for i in items), np.uint8).reshape(F, -1).copy()).to(dev)
        pn = torch.from_numpy(np.stack([i[1] for i in items])).to(dev)
        hp = torch.from_numpy(np.tile(self._hdr_pn_packed, (F, 1))).to(dev)
        band = torch.tensor([i[2] for i in items], dtype=torch.int32, device=dev)
        lo16 = torch.tensor([i[3] for i in items], dtype=torch.int32, device=dev)
        return tx_gpu.frames(pay, pn, hp, band, lo16, K=self.p.K).cpu().numpy()

    def _make_frames(self, nframes: int) -> np.ndarray:
        """Generate the next `nframes` frames and advance frame_ctr (the loop of rtwm/embedder.py:54-58)."""
        items = []
        for _ in range(nframes):
            items.append(self._frame_inputs(self.frame_ctr, self._build_payload()))
            self.frame_ctr = (self.frame_ctr + 1) % (2 ** 32)
        return self._synth(items)

    def _make_frame_chips(self) -> np.ndarray:
        """One frame for the current counter; does NOT advance frame_ctr (rtwm/embedder.py:78-151)."""
        return self._synth([self._frame_inputs(self.frame_ctr, self._build_payload())])[0]

    def _build_payload(self) -> bytes:
        """rtwm/embedder.py:153-168: "ESAL" | ctr_be32 | session nonce(8) | 11 random -> sealed 55 bytes."""
        meta = b"ESAL" + self.frame_ctr.to_bytes(4, "big") + self._session_nonce + secrets.token_bytes(11)
        assert len(meta) == 27
        blob = self.sec.seal(meta)
        assert len(blob) == 55
        return blob


class EmbedderBank:
    """Additive batched TX: S concurrent streams (one key, frame counter, session nonce and chip FIFO
    each) advanced in lock-step, one kernel launch per stage for all of them (SURVEY.md §8d config 5:
    4096 streams x 1024-sample blocks).  Per-stream semantics are those of WatermarkEmbedder.process
    (rtwm/embedder.py:44-75): frames are generated on demand, the mix level is computed per block.
    The host side (payload sealing, PN bits, hop bands) runs in the native threaded feeder."""

    def __init__(self, keys: list[bytes], params: TxParams | None = None, rand=None, prefetch: bool = False):
        """prefetch=True: the host side of the NEXT frames (payload sealing, PN, hop band: tx_prepare) runs in a
        background thread as soon as a block leaves fewer chips in the FIFO than one more block needs, so the
        block that does need them only launches the frame kernel (live use, TxService below)."""
        from .host_feeder import KeyBank
        import os
        self.p = params or TxParams()
        if self.p.N != N_DEFAULT or self.p.K != K_DEFAULT:
            raise ValueError("the B200 bank implements Polar(1024,448) frames")
        self.S = len(keys)
        self.bank = KeyBank(list(keys))
        self._rand = rand or os.urandom
        self.frame_ctr = np.zeros(self.S, np.uint32)
        self.session_nonce = np.frombuffer(self._rand(8 * self.S), np.uint8).reshape(self.S, 8).copy()
        self._fifo = None          # device tensor [S, avail]
        self._kidx = np.arange(self.S, dtype=np.int32)
        self._prefetch = bool(prefetch)
        self._pending = None       # (nframes, thread, result holder) of a running host preparation

    def make_frames(self, nframes: int) -> torch.Tensor:
        """Next `nframes` frames of every stream -> device float32 [S, nframes*1215]; advances frame_ctr."""
        dev = _dev()
        tx_gpu.set_filters(self.p.fs, self.p.preamble)
        S = self.S
        prep = None
        if self._pending is not None:
            pn, th, holder = self._pending
            th.join()
            self._pending = None
            if pn == nframes and "prep" in holder:
                prep = holder["prep"]
            elif "err" in holder:
                raise holder["err"]
        if prep is None:
            prep = self._prepare_host(nframes)
        chips = tx_gpu.frames(*(torch.from_numpy(prep[k]).to(dev, non_blocking=True)
                                for k in ("payload", "pn", "hdr_pn", "band", "ctr_lo16")), K=self.p.K)
        self.frame_ctr = ((self.frame_ctr.astype(np.uint64) + nframes) % (2 ** 32)).astype(np.uint32)
        return chips.view(S, nframes * FRAME_LEN)

    def _prepare_host(self, nframes: int):
        """host inputs of the next `nframes` frames of every stream (does not advance frame_ctr)"""
        F = self.S * nframes
        ctr = (self.frame_ctr[:, None].astype(np.uint64) + np.arange(nframes, dtype=np.uint64)[None, :]) % (2 ** 32)
        rnd = np.frombuffer(self._rand(23 * F), np.uint8).reshape(F, 23)
        return self.bank.tx_prepare(np.repeat(self._kidx, nframes), ctr.reshape(-1).astype(np.uint32),
                                    np.repeat(self.session_nonce, nframes, axis=0), rnd)

    def _start_prefetch(self, nframes: int):
        import threading
        holder = {}

        def work():
            try:
                holder["prep"] = self._prepare_host(nframes)
            except Exception as e:           # surfaced by the next make_frames
                holder["err"] = e
        th = threading.Thread(target=work, daemon=True)
        th.start()
        self._pending = (nframes, th, holder)

    def process(self, x):
        """x float32 [S, B] (numpy -> numpy, CUDA tensor -> CUDA tensor): one block of every stream."""
        dev = _dev()
        is_np = not isinstance(x, torch.Tensor)
        xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dev) if is_np else x.to(dev, torch.float32).contiguous()
        if xd.dim() != 2 or xd.shape[0] != self.S:
            raise ValueError(f"x must be [S={self.S}, B]")
        B = int(xd.shape[1])
        avail = 0 if self._fifo is None else int(self._fifo.shape[1])
        if avail < B:
            nf = (B - avail + FRAME_LEN - 1) // FRAME_LEN
            new = self.make_frames(nf)
            self._fifo = new if self._fifo is None else torch.cat([self._fifo, new], dim=1)
        chips = self._fifo[:, :B].contiguous()
        self._fifo = self._fifo[:, B:]
        if self._prefetch and self._pending is None and int(self._fifo.shape[1]) < B:
            self._start_prefetch((B - int(self._fifo.shape[1]) + FRAME_LEN - 1) // FRAME_LEN)
        out, _ = tx_gpu.mix(xd, chips, db_to_lin(self.p.target_rel_db), db_to_lin(self.p.floor_rel_dbfs))
        return out.cpu().numpy() if is_np else out


class TxService:
    """Live multi-stream embed service (rtwm/audioio.py:52-63 feeds WatermarkEmbedder.process one 1024-sample block
    per PortAudio callback; README: loop latency < 50 ms).  S streams advance in lock-step: `process_block` takes the
    next block of every stream in host memory and returns the watermarked block, through pinned staging buffers
    (double-buffered, so the returned array stays valid until the call after next) and one CUDA stream; the host
    crypto of upcoming frames runs in the background (EmbedderBank(prefetch=True)).  `latency_ms` holds the wall time
    of every call."""

    def __init__(self, keys: list[bytes], block: int = 1024, params: TxParams | None = None, rand=None):
        self.bank = EmbedderBank(keys, params, rand=rand, prefetch=True)
        self.S, self.B = len(keys), int(block)
        self._in = [torch.empty((self.S, self.B), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self._out = [torch.empty((self.S, self.B), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self._k = 0
        self._stream = torch.cuda.Stream()
        self.latency_ms: list[float] = []

    def process_block(self, x) -> np.ndarray:
        import time
        t0 = time.perf_counter()
        x = np.asarray(x, np.float32)
        if x.shape != (self.S, self.B):
            raise ValueError(f"block must be [S={self.S}, B={self.B}]")
        hin, hout = self._in[self._k], self._out[self._k]
        self._k ^= 1
        hin.numpy()[...] = x
        with torch.cuda.stream(self._stream):
            xd = hin.to(_dev(), non_blocking=True)
            y = self.bank.process(xd)
            hout.copy_(y, non_blocking=True)
        self._stream.synchronize()
        self.latency_ms.append(1e3 * (time.perf_counter() - t0))
        return hout.numpy()
