"""D2H bandwidth of one pinned copy vs the copy split over several streams (exploration for to_host)."""
import time, torch
n = 82 * 1024 * 1024 // 8
src = torch.randn(n, dtype=torch.float64, device="cuda")
dst = torch.empty(n, dtype=torch.float64, pin_memory=True)
for S in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(S)]
    chunks = list(zip(src.chunk(S), dst.chunk(S)))
    def run():
        ev = torch.cuda.Event(); ev.record()
        for st, (a, b) in zip(streams, chunks):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                b.copy_(a, non_blocking=True)
        for st in streams: st.synchronize()
    for _ in range(3): run()
    t0 = time.perf_counter()
    for _ in range(10): run()
    dt = (time.perf_counter() - t0) / 10
    print(f"S={S}: {dt*1e3:.2f} ms  {n*8/dt/1e9:.1f} GB/s")
