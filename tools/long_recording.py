"""configs[2]: one 1-hour 44.1 kHz synthetic recording with +-5 % time-scale and -15 dB SNR noise through
WatermarkDetector.verify (full +-200 counter fallback search, 400-try budget per band).  Prints one JSON line.
Under torchrun (WORLD_SIZE > 1) the four band scans are spread over the ranks
(echoseal_b200.sharding.verify_recording_sharded) and rank 0 prints the line, with the sequential verdict beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from echoseal_b200 import rx_gpu, embedder, detector, _native as N

def main():
    hours = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.05
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n48 = int(hours * 3600 * 48000)
    key = bench.bench_key(2024)
    g = torch.Generator(device=dev).manual_seed(2024)
    t0 = time.perf_counter()
    host = 0.05 * torch.randn((1, n48), device=dev, generator=g)
    tx = embedder.EmbedderBank([key])
    wm = tx.process(host)
    del host
    num, den = (21, 20) if scale > 1 else (19, 20)
    stretched = rx_gpu.resample(wm, den, num)                      # time-scale by num/den
    del wm
    a441 = rx_gpu.resample(stretched, 48000, 44100)
    del stretched
    p = float((a441.double() ** 2).mean())
    a441 = a441 + torch.randn(a441.shape, device=dev, generator=g) * np.sqrt(p * 10 ** 1.5)   # -15 dB SNR
    if world > 1:
        dist.broadcast(a441, src=0)              # the embedder draws fresh payload nonces: every rank must see rank 0's audio
    audio = a441[0].cpu().numpy()
    del a441
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    rx = detector.WatermarkDetector(key, list_size=8)
    N.KERNEL_TIMES = {}
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ok = rx.verify(audio, 44100)
    torch.cuda.synchronize(); dt = time.perf_counter() - t1
    kt = {k: float(sum(a.elapsed_time(b) for a, b in v)) for k, v in N.KERNEL_TIMES.items()}
    r = rx.last_result
    sharded = None
    if world > 1:
        from echoseal_b200.sharding import verify_recording_sharded
        rx2 = detector.WatermarkDetector(key, list_size=8)
        dist.barrier(); torch.cuda.synchronize(); t2 = time.perf_counter()
        ok2 = verify_recording_sharded(rx2, audio, 44100, device=dev)
        torch.cuda.synchronize(); dist.barrier(); dt2 = time.perf_counter() - t2
        sharded = {"world": world, "verdict": bool(ok2), "verify_seconds": dt2, "equals_sequential": bool(ok2) == bool(ok)}
        # the same recording split in TIME (global median / MAD through NCCL all-reduces of the histograms)
        from echoseal_b200.long_sharded import verify_recording_time_sharded
        rx3 = detector.WatermarkDetector(key, list_size=8)
        verify_recording_time_sharded(rx3, audio, 44100, device=dev)            # warm-up (allocations, NCCL)
        rx3 = detector.WatermarkDetector(key, list_size=8)
        dist.barrier(); torch.cuda.synchronize(); t3 = time.perf_counter()
        ok3 = verify_recording_time_sharded(rx3, audio, 44100, device=dev)
        torch.cuda.synchronize(); dist.barrier(); dt3 = time.perf_counter() - t3
        o = rx3.last_sharded
        same_peaks = all(list(o["peaks"][bi][:int(r.npeaks[bi])]) == [int(p) for p in r.peaks[bi][:int(r.npeaks[bi])]]
                         and int(o["npeaks"][bi]) == int(r.npeaks[bi]) for bi in range(4))
        sharded["time_sharded"] = {"verdict": bool(ok3), "verify_seconds": dt3, "equals_sequential": bool(ok3) == bool(ok),
                                   "sync_offsets_equal": bool(same_peaks), "scl_decodes": int(o["n_scl"]),
                                   "thr": [float(v) for v in o["thr"]]}
        if rank != 0:
            dist.destroy_process_group()
            return
    oracle = None
    if os.environ.get("ES_LONG_ORACLE") == "1":
        # the CPU oracle on the SAME 48 kHz samples the detector scanned (its own resampler output; the resampler has
        # its own parity test): thresholds, sync offsets, attempt lists, verdict at full size
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        from test_gpu_rx import compare_with_oracle
        a48 = rx._resample(audio, 44100)
        a48 = a48.cpu().numpy() if isinstance(a48, torch.Tensor) else np.asarray(a48)
        t4 = time.perf_counter()
        oracle = compare_with_oracle(r, np.ascontiguousarray(a48.reshape(-1), dtype=np.float32), key)
        oracle["oracle_seconds"] = time.perf_counter() - t4
        oracle["verdict_equal"] = oracle["oracle_verdict"] == bool(ok)
        oracle["scl_decodes_equal"] = oracle["oracle_scl_decodes"] == int(r.n_scl)
    print(json.dumps({"workload": f"configs[2]: {hours} h 44.1 kHz recording, time-scale x{num}/{den}, -15 dB SNR",
                      "samples_44k1": int(audio.size), "verdict": bool(ok), "verify_seconds": dt,
                      "audio_seconds_per_second": audio.size / 44100 / dt, "scl_decodes": int(r.n_scl),
                      "attempts_per_band": [len(a) for a in r.attempts], "npeaks": [int(v) for v in r.npeaks],
                      "thr": [float(v) for v in r.stats[:, 2]], "kernel_ms": kt, "generate_seconds": t_gen,
                      "band_sharded": sharded, "oracle_comparison": oracle}))
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
