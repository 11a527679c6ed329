"""Multi-GPU check of the domain-decomposed CMC / SA driver (run under torchrun, one rank per GPU):
every rank anneals the SAME lattice; the result must be identical on all ranks AND identical to a single-GPU run of the
same seed (computed on rank 0 with a second, unattached engine).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/domain_multi_gpu.py [f] [trials]"""
import hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from latticemontecarlo_b200 import capi, sharding, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
f = int(sys.argv[1]) if len(sys.argv) > 1 else 24
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 2000000
sa = (900.0, 100 * trials) if len(sys.argv) > 3 and sys.argv[3] == "sa" else None
js = "/tmp/coef_multi_%d.json" % rank
synth.write_synthetic_json(js)
occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)


def digest(e):
    st = e.cmc_state()
    return hashlib.sha256(e.get_occupancy(0).tobytes()).hexdigest()[:16] + " E=%.17g steps=%d acc=%d T=%.17g" % (
        st["energy"][0], st["steps"][0], st["accepted"][0], st["temperature"][0])


e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=1, device=local)
e.load_coefficients(js)
e.set_occupancy(occ)
sharding.attach_cmc_domain_peers(e, rank, world)
e.cmc_reset(*(sa or ()))
ms = []
for chunk in range(3):
    dist.barrier(); torch.cuda.synchronize()
    e.cmc_domain_run(trials, temperature=800.0, seed=5)
    ms.append(e.last_kernel_ms())
d = digest(e)
e0 = e.total_energy()
all_d = [None] * world
dist.all_gather_object(all_d, d)
st = e.cmc_state()
if rank == 0:
    ref = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=1, device=local)
    ref.load_coefficients(js)
    ref.set_occupancy(occ)
    e_start = ref.total_energy()
    ref.cmc_reset(*(sa or ()))
    ms1 = []
    for chunk in range(3):
        ref.cmc_domain_run(trials, temperature=800.0, seed=5)
        ms1.append(ref.last_kernel_ms())
    d1 = digest(ref)
    print("world %d f=%d: ranks identical %s, equals world-1 run %s" % (world, f, len(set(all_d)) == 1, d == d1))
    print("  multi:", d, "shape", e.cmc_domain_last_shape())
    print("  single:", d1, "shape", ref.cmc_domain_last_shape())
    print("  bookkeeping |(E_final - E_start) - energy| = %.2e" % abs((e0 - e_start) - st["energy"][0]))
    print("  kernel ms multi %s single %s -> speed-up %.2f, rate %.3e trials/s" % (["%.2f" % m for m in ms], ["%.2f" % m for m in ms1], sum(ms1) / sum(ms),
          3 * trials / (sum(ms) * 1e-3)))
dist.barrier()
dist.destroy_process_group()
