"""PCIe-only bound of the end-to-end step: the same host<->device copy pattern as bench.py's e2e pass (sizes per 64-frame
KITTI step), without any kernel.  Diagnostic only; prints ms/step for a few variants."""
import time
import torch

MB = 1 << 20
H2D_BIG = [29863424, 6758400, 3942400, 9984000]
H2D_SMALL = [131072, 256, 256, 256, 16384, 655360, 163840, 16640, 2304]
D2H = [3942400, 4505600, 256, 563200, 140800, 140800, 1024, 2048]


def run(h2d, d2h, nctx=4, steps=40):
    streams = [torch.cuda.Stream() for _ in range(nctx)]
    hin = [[torch.empty(n, dtype=torch.uint8).pin_memory() for n in h2d] for _ in range(nctx)]
    din = [[torch.empty(n, dtype=torch.uint8, device="cuda") for n in h2d] for _ in range(nctx)]
    hout = [[torch.empty(n, dtype=torch.uint8).pin_memory() for n in d2h] for _ in range(nctx)]
    dout = [[torch.empty(n, dtype=torch.uint8, device="cuda") for n in d2h] for _ in range(nctx)]

    def loop(k):
        for s in range(k):
            c = s % nctx
            if s >= nctx:
                streams[c].synchronize()
            with torch.cuda.stream(streams[c]):
                for h, d in zip(hin[c], din[c]):
                    d.copy_(h, non_blocking=True)
                for h, d in zip(hout[c], dout[c]):
                    h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
    loop(8)
    t0 = time.perf_counter()
    loop(steps)
    return (time.perf_counter() - t0) * 1e3 / steps


if __name__ == "__main__":
    tot = sum(H2D_BIG) + sum(H2D_SMALL)
    print("h2d bytes/step", tot, "d2h", sum(D2H))
    two = [sum(H2D_BIG[1:]) + sum(H2D_SMALL), H2D_BIG[0]]
    for name, a, b in [("h2d big only", H2D_BIG, []), ("h2d all", H2D_BIG + H2D_SMALL, []), ("h2d all + d2h", H2D_BIG + H2D_SMALL, D2H),
                       ("h2d one block + d2h", [tot], D2H), ("h2d one block + d2h one block", [tot], [sum(D2H)]),
                       ("h2d two blocks", two, []), ("h2d two blocks + d2h", two, D2H), ("h2d two blocks + d2h big only", two, [D2H[0], D2H[1], D2H[3]]),
                       ("h2d two blocks + d2h one block", two, [sum(D2H)])]:
        ms = run(a, b)
        print("%-32s %.3f ms/step  h2d %.1f GB/s  -> %.0f frames/s" % (name, ms, sum(a) / ms / 1e6, 64 / ms * 1e3))
