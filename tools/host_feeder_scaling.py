"""Thread scaling of the native host feeder (CPU only): tx_prepare (seal + PN + hop per frame) and KeyBank creation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from echoseal_b200.host_feeder import KeyBank

def main():
    S = 4096
    keys = [bench.bench_key(i) for i in range(S)]
    print("cpus", os.cpu_count())
    for nt in (1, 2, 4, 8, 16):
        t0 = time.perf_counter(); bank = KeyBank(keys, nthreads=nt); tk = time.perf_counter() - t0
        kidx = np.arange(S, dtype=np.int32); ctr = np.arange(S, dtype=np.uint32)
        sn = np.zeros((S, 8), np.uint8); rnd = np.zeros((S, 23), np.uint8)
        bank.tx_prepare(kidx, ctr, sn, rnd)
        t0 = time.perf_counter()
        for r in range(10):
            bank.tx_prepare(kidx, ctr + r + 1, sn, rnd)
        dt = (time.perf_counter() - t0) / 10
        print(f"threads {nt:2d}: KeyBank({S}) {1e3 * tk:7.2f} ms   tx_prepare({S} frames) {1e3 * dt:6.2f} ms")

if __name__ == "__main__":
    main()
