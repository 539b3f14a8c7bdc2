import torch, time
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return n * reps / (time.perf_counter() - t0) / 1e9
for _ in range(2):
    print("H2D only GB/s", round(run(True, False), 1), "D2H only", round(run(False, True), 1), "both (each dir)", round(run(True, True), 1))
