import sys, os, torch
sys.path.insert(0, os.getcwd())
import multimodal_learning_b200 as pkg
from multimodal_learning_b200 import _cabi
lib = _cabi.lib()
dev = torch.device("cuda:0")
B, cols, chunk = 1024, 16385, 2048
chunks = (cols + chunk - 1) // chunk
for world in (2, 8):
    rows_per = 2_000_000
    gen = torch.Generator(device=dev).manual_seed(0)
    cidx = torch.randint(0, rows_per * world, (B, cols), device=dev, generator=gen)
    counts = torch.zeros(world * B * chunks, dtype=torch.int32, device=dev)
    ids = torch.zeros(world * B * chunks * chunk, dtype=torch.int32, device=dev)
    def run():
        _cabi.check(lib.mml_shard_route_strided(_cabi.dptr(cidx, torch.int64), B, cols, chunk, rows_per, world,
                                                _cabi.dptr(counts), _cabi.dptr(ids), _cabi.cur_stream(dev)), "route")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    # exact check against a stable partition on the host for a few slots
    c = counts.view(world, B, chunks).cpu(); idv = ids.view(world, B, chunks, chunk).cpu(); h = cidx.cpu()
    ok = True
    for b in (0, 511, 1023):
        for ch in (0, 3, chunks - 1):
            seg = h[b, ch * chunk:min(cols, (ch + 1) * chunk)]
            for o in range(world):
                want = (seg[(seg // rows_per) == o] - o * rows_per).int()
                ok &= int(c[o, b, ch]) == want.numel() and torch.equal(idv[o, b, ch, :want.numel()], want)
    print(f"world {world}: route {ms*1e3:.1f} us, bytes {(B*cols*12)/ms/1e6:.0f} GB/s, exact {ok}")
