"""Multi-GPU parity check (run under torchrun): slab-decomposed realise == single-GPU realise."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fastbox_b200 import _lib  # noqa: E402
from fastbox_b200 import dist as fbd  # noqa: E402
from fastbox_b200 import kspace as ks  # noqa: E402


def tables(plan, N, L):
    pkf = lambda k: np.where(k > 0, 1e4 * (k / 0.02) / (1 + (k / 0.02) ** 2.5), 0.0)
    with np.errstate(all="ignore"):
        mode, tab, l0, dl = ks.choose_sqrt_pk_table(pkf, N, L, L, L, N ** 6. / L ** 3)
    plan.set_sqrt_pk(tab, mode, l0, dl)
    fn = lambda kp, kl: (1. - np.exp(-0.5 * (kl / 0.001) ** 2.)) * np.exp(-0.5 * (kp / 0.1) ** 2.)
    ft = ks.filter_tables(fn, N, L, L, L)
    plan.set_filter(ft.tperp, ft.tpar, ft.tdense)
    edges = ks.pk_bin_edges(2 * np.pi / L, 2 * np.pi * np.sqrt(3.) * N / L, 50)
    plan.set_pk_bins(ks.bin_thresholds(edges))


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="p2p", choices=["p2p", "nccl"],
                    help="p2p: exchange inside the library (peer stores over NVLink); nccl: all_to_all_single")
    ap.add_argument("--sizes", default="64,256")
    args = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
    ok = True
    for N in [int(x) for x in args.sizes.split(",")]:
        L = 2000.0 * N / 1024
        chunks = 2 if (N // 2) // world % 2 == 0 else 1
        flags = _lib.F_SQRTPK | _lib.F_FILTER
        if args.mode == "p2p":
            dr = fbd.NvlinkRealiser(N, (L, L, L), rank, world, lr, chunks=chunks)
            tables(dr.plan, N, L)
            for rep in range(3):                       # repeated steps: both receive buffers, rising epochs
                fbuf, pk, sums = dr.realise(7, flags, want_pk=True)
            field = torch.from_numpy(dr.field_host()).cuda()
        else:
            eng = fbd.CudaEngine(N, (L, L, L), rank, world, lr, chunks=chunks)
            tables(eng.plan, N, L)
            dr = fbd.DistributedRealiser(eng)
            if chunks > 1:
                field, pk, sums = dr.realise_overlapped(7, flags, want_pk=True)
            else:
                field, pk, sums = dr.realise(7, flags, want_pk=True)
        torch.cuda.synchronize()
        gathered = [torch.empty_like(field) for _ in range(world)]
        dist.all_gather(gathered, field)
        if rank == 0:
            full = torch.cat(gathered, dim=1).cpu().numpy()            # slabs along y
            plan = _lib.Plan(N, L, L, L, lr)
            tables(plan, N, L)
            ref = np.empty((N, N, N), np.float32)
            res, _ = plan.realise(None, None, seed=7, flags=flags, field_out=ref, want_pk=True)
            err = np.linalg.norm(full.astype(np.float64) - ref) / np.linalg.norm(ref)
            same_counts = np.array_equal(pk["count"], res["count"])
            pk_err = np.nanmax(np.abs(pk["sum1"] - res["sum1"]) / np.maximum(np.abs(res["sum1"]), 1e-300))
            print("[%s] N=%d world=%d field rel-L2 %.2e (bit identical %s), counts equal %s, sum1 rel err %.1e" %
                  (args.mode, N, world, err, np.array_equal(full, ref), same_counts, pk_err), flush=True)
            ok = ok and err < 1e-6 and same_counts and pk_err < 1e-10
            fwd_ref = plan.field_to_spectrum(ref, want_pk=True)
            plan.close()
        # forward direction: P(k) of the sharded field through the reverse all-to-all
        pk_f = dr.power_spectrum()
        if rank == 0:
            same = np.array_equal(pk_f["count"], fwd_ref["count"])
            e1 = np.nanmax(np.abs(pk_f["sum1"] - fwd_ref["sum1"]) / np.maximum(np.abs(fwd_ref["sum1"]), 1e-300))
            print("      forward P(k): counts equal %s, sum1 rel err %.1e" % (same, e1), flush=True)
            ok = ok and same and e1 < 1e-10
        dist.barrier()
    if rank == 0:
        print("DIST CHECK", "OK" if ok else "FAILED", flush=True)
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
