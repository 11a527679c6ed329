"""Multi-GPU single-lattice CMC check + timing.  Launch one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
        tools/cmc_multi_gpu.py <factor> [trials]
Every rank runs lmc_cmc_grid_run on its own full replica of the lattice with the evaluation sharded over the ranks
(exchange through peer memory inside the kernel); rank 0 then repeats the run alone (world = 1) on a second engine and
checks that occupancy, energy, step and accept counters are IDENTICAL."""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latticemontecarlo_b200 import capi, sharding, synth  # noqa: E402


def main():
    f = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = tempfile.mkdtemp()
    js = os.path.join(tmp, "c.json")
    synth.write_synthetic_json(js)
    occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)
    e = capi.Engine(f, n_walkers=1, device=local)
    e.load_coefficients(js)
    e.set_occupancy(occ)
    e0 = e.total_energy()
    sharding.attach_cmc_peers(e, rank, world)
    results = {}
    for label, sa in (("cmc", None), ("sa", (900.0, 4 * trials))):
        e.set_occupancy(occ)
        e.cmc_reset(*(sa or ()))
        dist.barrier(); torch.cuda.synchronize()
        e.cmc_grid_run(trials // 10, temperature=800.0, seed=5)
        rates = []
        for _ in range(3):
            dist.barrier(); torch.cuda.synchronize()
            s0 = e.cmc_state()
            e.cmc_grid_run(trials, temperature=800.0, seed=5)
            ms = sharding.max_over_ranks([e.last_kernel_ms()], device="cuda")[0]
            rates.append(float(e.cmc_state()["steps"][0] - s0["steps"][0]) / ms * 1e3)
        st = e.cmc_state()
        final = e.get_occupancy(0)
        digest = hashlib.sha256(final.tobytes()).hexdigest()
        digests = [None] * world
        dist.all_gather_object(digests, (digest, float(st["energy"][0]), int(st["steps"][0]), int(st["accepted"][0]), float(st["temperature"][0])))
        results[label] = (rates, digests, abs((e.total_energy() - e0) - st["energy"][0]))
    if rank == 0:
        solo = capi.Engine(f, n_walkers=1, device=local)
        solo.load_coefficients(js)
        for label, sa in (("cmc", None), ("sa", (900.0, 4 * trials))):
            solo.set_occupancy(occ)
            solo.cmc_reset(*(sa or ()))
            solo.cmc_grid_run(trials // 10, temperature=800.0, seed=5)
            ms = []
            for _ in range(3):
                s0 = solo.cmc_state(); solo.cmc_grid_run(trials, temperature=800.0, seed=5)
                ms.append(float(solo.cmc_state()["steps"][0] - s0["steps"][0]) / solo.last_kernel_ms() * 1e3)
            st = solo.cmc_state()
            ref = (hashlib.sha256(solo.get_occupancy(0).tobytes()).hexdigest(), float(st["energy"][0]), int(st["steps"][0]), int(st["accepted"][0]),
                   float(st["temperature"][0]))
            rates, digests, book = results[label]
            same = all(d == ref for d in digests)
            print("%s f=%d world=%d: identical_to_world1=%s bookkeeping=%.2e steps=%d accepted=%d | rate world=%d %s | world=1 %s"
                  % (label, f, world, same, book, ref[2], ref[3], world, ["%.3g" % r for r in rates], ["%.3g" % r for r in ms]), flush=True)
            if not same:
                print("  MISMATCH", ref, digests, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
