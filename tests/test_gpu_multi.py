"""World-size-2 run of the product path on the GPU box: the multistart guesses of Optimize.optimal
and the rows / cells of history matching are block-partitioned over two ranks (gloo process group,
both ranks on cuda:0 -- the driver's GPU test box has one GPU; under torchrun on a multi-GPU box the
same code runs one rank per GPU over NCCL, which bench.py --gpus N exercises)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import contextlib, io, os, shutil, sys
import numpy as np
import torch.distributed as dist
ROOT, port, rank, work, backend = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
sys.path.insert(0, ROOT)
if backend == "nccl":                 # one rank per GPU, cell statistics / guess tables travel over NCCL
    import torch
    os.environ["GPE_DEVICE"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + port, rank=rank, world_size=2,
                            device_id=torch.device("cuda", rank))
else:                                 # single-GPU box: both ranks on cuda:0, gloo for the exchanges
    os.environ["GPE_DEVICE"] = "0"
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + port, rank=rank, world_size=2)
import gp_emu_uqsa_b200 as g
import gp_emu_uqsa_b200.history_match as h
gdir = os.path.join(ROOT, "tests", "golden")
wd = os.path.join(work, "r%d" % rank)
os.makedirs(wd)
os.chdir(wd)
for f in os.listdir(os.path.join(gdir, "toy-sim")):
    shutil.copy(os.path.join(gdir, "toy-sim", f), wd)
gold = np.load(os.path.join(gdir, "toysim.npz"))
with contextlib.redirect_stdout(io.StringIO()):
    # only rank 0 is seeded like the golden run; rank 1 starts from another generator state (as an unseeded torchrun job
    # would) and must still shuffle, guess and design exactly like rank 0 (_dist.sync_numpy_rng)
    np.random.seed(0 if rank == 0 else 4242)
    E = g.setup("toy-sim_config")
    g.train(E)
assert E.opt_T.last_evals > 0                         # this rank optimised its own block of the 10 guesses
assert np.allclose(E.par.delta, gold["delta"], rtol=2e-4) and abs(E.par.sigma - gold["sigma"]) < 2e-4 * gold["sigma"]
state = np.concatenate([E.par.delta, [E.par.sigma], E.par.beta])
both = [None, None]
dist.all_gather_object(both, state)
assert np.array_equal(both[0], both[1]), "ranks disagree on the selected optimum"
# history matching: rows of nonimp_data and cells of imp_plot split over the ranks
hm = np.load(os.path.join(gdir, "hmapi_n100_d3.npz"))
emuls = []
with contextlib.redirect_stdout(io.StringIO()):
    for o in (0, 1):
        for fn in ("hm%d_config_r" % o, "hm%d_beliefs-0f" % o, "hm%d_input-o0-0f" % o, "hm%d_output-o0-0f" % o):
            open(fn, "wb").write(bytes(hm["file_" + fn]))
        emuls.append(g.setup("hm%d_config_r" % o, datashuffle=False, scaleinputs=True))
    zs, ve, cm = list(hm["zs"]), list(hm["var_extra"]), float(hm["cm"])
    np.savetxt("sim_in", hm["sim_in"], fmt="%.17g")
    np.savetxt("sim_out", np.column_stack([hm["sim_in"][:, 0], hm["sim_in"][:, 1]]), fmt="%.17g")
    cnt = h.nonimp_data(emuls, zs, cm, ve, ["sim_in", "sim_out"], maxno=1)
    np.random.seed(77 if rank == 0 else 5)
    h.imp_plot(emuls, zs, cm, ve, maxno=2, olhcmult=30, grid=4, plot=False, fileStr="g")
assert cnt == int(hm["nonimp_count"])
dist.barrier()
if rank == 0:
    assert np.allclose(np.atleast_2d(np.loadtxt("nonimp_sim_in")), hm["nonimp_in"], rtol=0, atol=1e-15)
    for s_ in ([0, 1], [0, 2], [1, 2]):
        for m in (1, 2):
            assert np.allclose(np.loadtxt("g_%d_IMP_%d_%d" % (m, s_[0], s_[1])), hm["g_%d_IMP_%d_%d" % (m, s_[0], s_[1])], rtol=1e-7, atol=1e-9)
            assert np.array_equal(np.loadtxt("g_%d_ODP_%d_%d" % (m, s_[0], s_[1])), hm["g_%d_ODP_%d_%d" % (m, s_[0], s_[1])])
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_ranks_shard_multistart_and_history_match(tmp_path):
    import torch
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), str(tmp_path), backend], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in (0, 1)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d ok" % r) in o, o[-3000:]
