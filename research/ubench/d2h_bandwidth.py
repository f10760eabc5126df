import torch, time
torch.cuda.set_device(0)
n = 2 * 1024**3 // 8
d = torch.empty(n, dtype=torch.float64, device='cuda').normal_()
h = torch.empty(n, dtype=torch.float64).pin_memory()
def run(nstreams, chunk_mib):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    ce = chunk_mib * 1024 * 1024 // 8
    torch.cuda.synchronize(); t = time.time()
    i = 0; k = 0
    while i < n:
        j = min(n, i + ce)
        with torch.cuda.stream(streams[k % nstreams]):
            h[i:j].copy_(d[i:j], non_blocking=True)
        i = j; k += 1
    torch.cuda.synchronize()
    return n * 8 / (time.time() - t) / 1e9
for ns, cm in ((1, 2048), (1, 256), (2, 256), (4, 64), (1, 32), (2, 32)):
    run(ns, cm)
    print("streams %d chunk %4d MiB: %.1f GB/s" % (ns, cm, max(run(ns, cm) for _ in range(3))))
