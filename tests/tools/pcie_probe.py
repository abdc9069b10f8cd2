"""PCIe probe for the end-to-end path: H2D / D2H alone and overlapped, whole and in pieces (pinned host memory)."""
import sys, time, torch
GB = 1 << 30
up_n, dn_n = int(3.13e9), int(5.78e9)
dev = torch.device("cuda", 0)
h_up = torch.empty(up_n, dtype=torch.uint8).pin_memory(); h_up.fill_(65)
h_dn = torch.empty(dn_n, dtype=torch.uint8).pin_memory(); h_dn.fill_(1)
d_up = torch.empty(up_n, dtype=torch.uint8, device=dev)
d_dn = torch.zeros(dn_n, dtype=torch.uint8, device=dev)
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best

def up_whole():
    with torch.cuda.stream(s_up): d_up.copy_(h_up, non_blocking=True)
def dn_whole():
    with torch.cuda.stream(s_dn): h_dn.copy_(d_dn, non_blocking=True)
def up_pieces(piece):
    with torch.cuda.stream(s_up):
        for o in range(0, up_n, piece): d_up[o:o + piece].copy_(h_up[o:o + piece], non_blocking=True)
def dn_pieces(piece):
    with torch.cuda.stream(s_dn):
        for o in range(0, dn_n, piece): h_dn[o:o + piece].copy_(d_dn[o:o + piece], non_blocking=True)

t = timed(up_whole); print(f"H2D whole      {t*1e3:7.1f} ms  {up_n/t/1e9:6.1f} GB/s")
t = timed(dn_whole); print(f"D2H whole      {t*1e3:7.1f} ms  {dn_n/t/1e9:6.1f} GB/s")
for piece in (4 << 20, 32 << 20, 256 << 20):
    t = timed(lambda: up_pieces(piece)); print(f"H2D {piece>>20:4d} MB pcs {t*1e3:7.1f} ms  {up_n/t/1e9:6.1f} GB/s")
    t = timed(lambda: dn_pieces(piece)); print(f"D2H {piece>>20:4d} MB pcs {t*1e3:7.1f} ms  {dn_n/t/1e9:6.1f} GB/s")
t = timed(lambda: (up_whole(), dn_whole())); print(f"both whole     {t*1e3:7.1f} ms  {(up_n+dn_n)/t/1e9:6.1f} GB/s combined")
t = timed(lambda: (up_pieces(32 << 20), dn_pieces(32 << 20))); print(f"both 32 MB pcs {t*1e3:7.1f} ms  {(up_n+dn_n)/t/1e9:6.1f} GB/s combined")
# delayed download: starts when 10% of the upload is done (what a streaming pipeline would do)
def staggered():
    ev = torch.cuda.Event()
    with torch.cuda.stream(s_up):
        k = up_n // 10
        d_up[:k].copy_(h_up[:k], non_blocking=True); ev.record()
        d_up[k:].copy_(h_up[k:], non_blocking=True)
    with torch.cuda.stream(s_dn):
        s_dn.wait_event(ev)
        h_dn.copy_(d_dn, non_blocking=True)
t = timed(staggered); print(f"staggered      {t*1e3:7.1f} ms")
