"""CPU, world_size 2 (gloo): the one-process-per-GPU row-band path of bench.py / fixca.bands.

No kernel runs here (the product has no CPU path): each rank plans its band with the library's own
host logic, keeps ONLY the source rows the plan says it needs (everything else is poisoned), lets the
oracle stand in for the GPU on that band, and the bands are gathered on rank 0 and compared with
the oracle's full-image result.  This pins: band split, halo sufficiency, row bookkeeping
(src_row0 / dst_row0 offsets) and the gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    # h, w, ch, dtype, interpolation, params
    (203, 157, 3, "u2", 2, dict(blue=3.0, red=-2.0, lens_x=78, lens_y=101, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)),
    (97, 64, 4, "u1", 1, dict(blue=-6.0, red=2.4, lens_x=0, lens_y=0)),
    (120, 90, 3, "f4", 0, dict(blue=30.0, red=-30.0, lens_x=45, lens_y=60, y_blue=30.0, y_red=-30.0)),
    (5, 40, 3, "u1", 2, dict(blue=1.0, red=-1.5, lens_x=20, lens_y=2)),          # fewer rows than a halo
]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    for p in (os.path.join(ROOT, "gimp-fix-ca_b200"), os.path.join(ROOT, "oracle"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import fixca
    import oracle as orc
    from fixca import bands

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chk = orc.best_checker()
        ok = True
        for n, (h, w, ch, dt, interp, kw) in enumerate(CASES):
            img = orc.synth_image(h, w, ch, dt, seed=900 + n)
            P = fixca.FixCaParams(interpolation=interp, **kw)
            plan = bands.plan_band(w, h, P, rank, world)
            # the rank's view of the image: its rows + halo are real, the rest is poison
            poison = np.full_like(img, 0x5A if img.dtype.kind != "f" else np.nan)
            mine = poison.copy()
            if plan.src_rows > 0:
                mine[plan.src_lo:plan.src_hi + 1] = img[plan.src_lo:plan.src_hi + 1]
            out = np.zeros_like(img)
            if plan.y2 > plan.y1:
                chk.region(mine, orc.Params(interpolation=interp, **kw), plan.y1, plan.y2, dst=out)
            band = torch.from_numpy(out[plan.y1:plan.y2].copy())
            full = bands.gather_bands(band, plan, dst_rank=0)
            if rank == 0:
                want = chk.region(img, orc.Params(interpolation=interp, **kw))
                got = full.numpy()
                same = got.shape == want.shape and got.tobytes() == want.tobytes()
                ok = ok and same
                if not same:
                    out_q.put(("mismatch", n))
            else:
                assert full is None
            # every rank agrees on the global plan
            t = torch.tensor([plan.y1, plan.y2], dtype=torch.int64)
            allp = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allp, t)
            edges = [int(v) for p_ in allp for v in p_]
            assert edges[0] == 0 and edges[-1] == h and all(edges[2 * i + 1] == edges[2 * i + 2] for i in range(world - 1))
        if rank == 0:
            out_q.put(("ok", ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_row_bands_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    results = []
    while not q.empty():
        results.append(q.get())
    assert ("ok", True) in results, results


def test_band_plans_cover_the_image_without_overlap():
    for p in (os.path.join(ROOT, "gimp-fix-ca_b200"),):
        if p not in sys.path:
            sys.path.insert(0, p)
    import fixca
    from fixca import bands

    P = fixca.FixCaParams(interpolation=2, blue=3.0, red=-2.0, lens_x=6144, lens_y=4096 * 8, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)
    for world in (1, 2, 4, 8):
        h = 8192 * world
        plans = [bands.plan_band(12288, h, P, r, world) for r in range(world)]
        assert plans[0].y1 == 0 and plans[-1].y2 == h
        for a, b in zip(plans, plans[1:]):
            assert a.y2 == b.y1
        for pl in plans:
            assert pl.y2 - pl.y1 == 8192                       # weak scaling: equal bands
            assert pl.src_lo <= pl.y1 and pl.src_hi >= pl.y2 - 1
            assert pl.halo_rows <= 2 * 64                      # halo <= ~64 rows per side


def _peer_worker(rank, world, port, out_q, one_gpu=False):
    """One process per GPU (NCCL): every rank's kernel stores its band into rank 0's frame (PeerFrame, CUDA IPC
    peer mapping over NVLink); rank 0 compares the frame with the oracle's full image.  one_gpu: the same
    processes on ONE device (gloo for the rendezvous; the frame is still another process's allocation, mapped
    through the same fixca_cuda_frame_open) -- what a single-GPU box can check of this path."""
    for p in (os.path.join(ROOT, "gimp-fix-ca_b200"), os.path.join(ROOT, "oracle"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import fixca
    import oracle as orc
    from fixca import bands

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0 if one_gpu else rank)
    dev = torch.device("cuda", 0 if one_gpu else rank)
    if one_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        chk = orc.best_checker()
        ok = True
        cases = [
            (1203, 1157, 3, "u2", 2, fixca.PRECISION_FAST, 1, dict(blue=3.0, red=-2.0, lens_x=578, lens_y=601, x_blue=0.7, x_red=-0.4, y_blue=0.3, y_red=-0.9)),
            (997, 640, 4, "u1", 1, fixca.PRECISION_EXACT, 0, dict(blue=-6.0, red=2.4, lens_x=0, lens_y=0)),
            (1200, 900, 3, "f4", 0, fixca.PRECISION_EXACT, 0, dict(blue=30.0, red=-30.0, lens_x=450, lens_y=600, y_blue=30.0, y_red=-30.0)),
        ]
        for n, (h, w, ch, dt, interp, flags, tol, kw) in enumerate(cases):
            flags |= fixca.PADDING_SCRATCH          # the frames are pitched: the bytes after a row's end are scratch
            img = orc.synth_image(h, w, ch, dt, seed=700 + n)
            P = fixca.FixCaParams(interpolation=interp, **kw)
            plan = bands.plan_band(w, h, P, rank, world)
            bpp = ch * img.dtype.itemsize
            row_bytes = w * bpp
            pitch = (row_bytes + 127) // 128 * 128
            # this rank holds ONLY its band + halo rows
            mine = np.zeros((plan.src_rows, pitch), dtype=np.uint8)
            mine[:, :row_bytes] = img[plan.src_lo:plan.src_hi + 1].view(np.uint8).reshape(plan.src_rows, row_bytes)
            d_src = torch.from_numpy(mine).to(dev)
            frame = bands.PeerFrame(h, pitch, owner=0)
            if rank == 0:
                frame.as_tensor().fill_(0x5A)
            frame.sync()
            bands.run_band_into_frame(plan, d_src.data_ptr(), pitch, frame, bpp, fixca.bpc_of(img.dtype), P, flags,
                                      torch.cuda.current_stream().cuda_stream)
            frame.sync()
            want = chk.region(img, orc.Params(interpolation=interp, **kw))

            def matches(t):
                got = t[:, :row_bytes].cpu().numpy().copy().view(img.dtype).reshape(h, w, ch)
                if tol == 0:
                    return got.tobytes() == want.tobytes()
                return int(np.abs(got.astype(np.int64) - want.astype(np.int64)).max()) <= tol

            if rank == 0:
                same = matches(frame.as_tensor())
                ok = ok and same
                if not same:
                    out_q.put(("mismatch", n))
            frame.close()
            # all-gather form: every rank owns a frame, every rank's kernel stores its band into all of them
            allf = bands.AllFrames(h, pitch)
            allf.as_tensor().fill_(0x5A)
            allf.sync()
            bands.run_band_into_all(plan, d_src.data_ptr(), pitch, allf, bpp, fixca.bpc_of(img.dtype), P, flags,
                                    torch.cuda.current_stream().cuda_stream)
            allf.sync()
            same = matches(allf.as_tensor())
            if not same:
                out_q.put(("all-gather mismatch", n, rank))
            flag = torch.tensor([1 if same else 0])
            if not one_gpu:
                flag = flag.to(dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = ok and bool(flag.item())
            allf.close()
        if rank == 0:
            out_q.put(("ok", ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_bands_stored_into_another_process_frame_one_gpu():
    """The peer-frame path on a single-GPU box: two processes on device 0, rank 1's kernel stores its band into the frame
    rank 0 allocated (CUDA IPC mapping), rank 0 checks the assembled frame against the oracle.  (The kernels of the two
    processes do not wait on one another.)"""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q, True)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    results = []
    while not q.empty():
        results.append(q.get())
    assert ("ok", True) in results, results


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_bands_stored_into_a_peer_frame_world2_nccl():
    """Needs two GPUs of one box (skipped on a single-GPU box): compute + gather in one kernel over NVLink."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    results = []
    while not q.empty():
        results.append(q.get())
    assert ("ok", True) in results, results
