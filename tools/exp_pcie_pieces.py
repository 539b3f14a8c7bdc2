#!/usr/bin/env python3
"""Both-direction pinned copy rate against the piece size (what a chunked pipeline can reach at best)."""
import time, torch
tot = 256 << 20
h_in = torch.empty(tot, dtype=torch.uint8).pin_memory(); h_out = torch.empty(tot, dtype=torch.uint8).pin_memory()
d_in = torch.empty(tot, dtype=torch.uint8, device="cuda"); d_out = torch.empty(tot, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for mib in (1, 2, 4, 8, 16, 64, 256):
    n = mib << 20
    def run():
        for o in range(0, tot, n):
            with torch.cuda.stream(s1): d_in[o:o + n].copy_(h_in[o:o + n], non_blocking=True)
            with torch.cuda.stream(s2): h_out[o:o + n].copy_(d_out[o:o + n], non_blocking=True)
    run(); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(); run(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    print("%4d MiB pieces: %.2f GB/s per direction" % (mib, tot / dt / 1e9))
