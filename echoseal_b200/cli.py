"""Batch front end of the B200 hot path — the "next" row SURVEY.md §8f-3: the reference's `rx_app.py`
verifies ONE file per process (rx_app.py:21-29); this front end verifies a whole directory / file list in
one batched pass per GPU and prints one verdict per file in the reference's wording.  `embed` is the
offline counterpart of `tx_app.py` (file in -> watermarked file out; no live audio device here).

    python -m echoseal_b200.cli verify --key <64 hex | keyfile> a.wav b.npy dir/
    python -m echoseal_b200.cli embed  --key <64 hex | keyfile> in.wav out.wav

Audio files: PCM16/PCM32/float32 .wav (stdlib `wave`, mono or first channel) and .npy arrays + `--fs`."""
from __future__ import annotations
import argparse
import os
import sys
import wave
import numpy as np


def load_key(path_or_hex: str) -> bytes:
    """rx_app.py:15-19"""
    stripped = path_or_hex.strip()
    if len(stripped) in (32, 48, 64) and all(c in "0123456789abcdefABCDEF" for c in stripped):
        return bytes.fromhex(stripped)
    return open(stripped, "rb").read()


def read_audio(path: str, fs_default: int):
    if path.endswith(".npy"):
        a = np.load(path)
        return np.asarray(a.reshape(a.shape[0], -1)[:, 0] if a.ndim > 1 else a), fs_default
    with wave.open(path, "rb") as w:
        fs, nch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if sw == 2:
        a = np.frombuffer(raw, "<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        a = np.frombuffer(raw, "<i4").astype(np.float32) / 2147483648.0
    else:
        raise SystemExit(f"{path}: unsupported sample width {sw}")
    return a.reshape(-1, nch)[:, 0].copy(), fs


def write_wav(path: str, x: np.ndarray, fs: int):
    if path.endswith(".npy"):
        np.save(path, x.astype(np.float32))
        return
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(fs)
        w.writeframes((np.clip(x, -1.0, 1.0) * 32767.0).astype("<i2").tobytes())


def expand(paths):
    out = []
    for p in paths:
        if os.path.isdir(p):
            out += sorted(os.path.join(p, f) for f in os.listdir(p) if f.endswith((".wav", ".npy")))
        else:
            out.append(p)
    return out


def cmd_verify(args) -> int:
    from .detector import verify_batch
    from . import rx_gpu
    import torch
    key = load_key(args.key)
    if len(key) != 32:
        raise SystemExit("❌  Key must be 256-bit (64 hex chars).")
    files = expand(args.audio)
    clips = [read_audio(f, args.fs) for f in files]
    dev = torch.device("cuda", torch.cuda.current_device())
    # group by (sample rate, length): each group is one batched pass
    groups = {}
    for i, (a, fs) in enumerate(clips):
        a = np.asarray(a, np.float32)
        if fs != 48_000 and a.size:
            a = rx_gpu.resample(torch.from_numpy(a[None]).to(dev), fs, 48_000)[0].cpu().numpy()
        groups.setdefault(a.size, []).append((i, a))
    verdict = [False] * len(files)
    for n, items in groups.items():
        batch = np.stack([a for _, a in items]) if n else np.zeros((len(items), 0), np.float32)
        v = verify_batch([key] * len(items), batch, list_size=8)
        for (i, _), ok in zip(items, v):
            verdict[i] = bool(ok)
    for f, ok in zip(files, verdict):
        print(f"{f}: " + ("✅  authentic" if ok else "⚠️  tampered / no watermark"))      # rx_app.py:29
    return 0


def cmd_embed(args) -> int:
    from .embedder import WatermarkEmbedder
    key = load_key(args.key)
    if len(key) != 32:
        raise SystemExit("❌  Key must be 256-bit (64 hex chars).")
    a, fs = read_audio(args.src, args.fs)
    if fs != 48_000:
        raise SystemExit("embed: input must be 48 kHz (TxParams.fs, rtwm/embedder.py:21)")
    tx = WatermarkEmbedder(key)
    out = np.concatenate([tx.process(a[i:i + args.block]) for i in range(0, a.size, args.block)]) if a.size else a
    write_wav(args.dst, out, fs)
    print(f"{args.dst}: {tx.frame_ctr} frames embedded", file=sys.stderr)
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="echoseal-b200", description="Batched EchoSeal verify / embed on B200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    v = sub.add_parser("verify", help="verify many files in one batched GPU pass")
    v.add_argument("--key", required=True, help="256-bit hex key (64 hex chars) or path to keyfile")
    v.add_argument("--fs", type=int, default=48_000, help="sample rate of .npy inputs")
    v.add_argument("audio", nargs="+", help="files or directories (.wav / .npy)")
    v.set_defaults(fn=cmd_verify)
    e = sub.add_parser("embed", help="watermark one file")
    e.add_argument("--key", required=True)
    e.add_argument("--fs", type=int, default=48_000)
    e.add_argument("--block", type=int, default=1024, help="process() block size (rtwm/audioio.py:18)")
    e.add_argument("src"); e.add_argument("dst")
    e.set_defaults(fn=cmd_embed)
    args = ap.parse_args(argv)
    return args.fn(args)


if __name__ == "__main__":
    raise SystemExit(main())
