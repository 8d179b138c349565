"""CPU-only tests: the C-ABI library loads and exports what include/diffiqult_b200.h declares, the host-side logic
(system flattening, index maps, work decomposition, multi-rank plumbing over gloo) behaves, and the product refuses
to compute without a GPU (no CPU fallback)."""
import ctypes as C
import itertools
import os
import re
import socket
import sys

import numpy as np
import pytest

import systems
from diffiqult_b200 import System_mol, Tasks, Tools, _capi, synthetic
from diffiqult_b200.Basis import basis_set_3G_STO, basis_set_2G_STO_uncontracted
from oracle import cpu_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "diffiqult_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dq_[a-z0-9_]+)\s*\(", hdr)) - {"dq_allreduce_fn"})
    assert len(declared) >= 12
    lib = _capi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_capi.EXPORTS) == declared


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails loudly with status -3."""
    if _capi.lib().dq_device_count() > 0:
        pytest.skip("a GPU is present")
    mol, basis, ne = systems.SYSTEMS["h2_sto3g"]
    s = System_mol(mol, basis, ne)
    with pytest.raises(_capi.DiffiqultB200Error, match="no CUDA device"):
        Tasks(s, "/tmp/dq_nogpu").runtask("Energy")
    w, V = np.zeros(2), np.zeros(4)
    assert _capi.lib().dq_eigh(C.c_int32(2), C.c_int32(1), np.eye(2).ctypes, w.ctypes, V.ctypes, C.c_int32(0)) == -3


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "diffiqult_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|from\s+\.+oracle|cpu_oracle|rhf_oracle|oracle/|fwdmode|ref_loader", txt), f


def test_index_maps_match_reference_formulas():
    for n in (1, 2, 5, 9):
        m = n * (n + 1) // 2
        assert Tools.eri_vector_size(n) == m * (m + 1) // 2
        seen = set()
        for i, j, k, l in itertools.product(range(n), repeat=4):
            idx = Tools.eri_index(i, j, k, l, n)
            assert idx == orc.eri_index(i, j, k, l, n)
            seen.add(idx)
        assert seen == set(range(Tools.eri_vector_size(n)))
        for x in range(m):
            i, j = Tools.vec_tomatrix(x, n)
            assert i <= j and Tools.matrix_tovector(i, j, n) == x


def test_system_mol_layout_and_angstrom_quirk():
    s = System_mol(synthetic.H2O, basis_set_3G_STO, 10, "h2o")
    assert s.nbasis == 7 and len(s.alpha) == 21 and s.ne == 5 and s.natoms == 3
    assert s.l == [(0, 0, 0), (0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0), (0, 0, 0)]
    assert list(s.list_contr) == [3] * 7
    # p components carry their own copies of the primitives (Molecule.py:60-71)
    assert np.array_equal(s.alpha[6:9], s.alpha[9:12]) and np.array_equal(s.alpha[9:12], s.alpha[12:15])
    a = System_mol(synthetic.H2O, basis_set_3G_STO, 10, "h2o", angs=True)
    assert np.allclose(a.xyz, 0.529177249 * s.xyz)   # sic: Molecule.py:185-188 multiplies
    u = System_mol(synthetic.H2, basis_set_2G_STO_uncontracted, 2)
    assert u.nbasis == 4 and list(u.list_contr) == [1, 1, 1, 1]
    # reference quirk kept: shifted=True never fills list_contr (Molecule.py:73-79), so nbasis = 0 and the reshape of
    # the centres at Molecule.py:171 raises, exactly as in the reference
    with pytest.raises(ValueError):
        System_mol([], [[(0, 0, 0), 1.0, 1.0, (0.0, 0.0, 0.0)]], 2, shifted=True)


def test_only_s_and_p_functions():
    with pytest.raises(NotImplementedError):
        _capi.ltypes([(2, 0, 0)])
    with pytest.raises(NotImplementedError):
        _capi.ltypes([(1, 1, 0)])
    assert list(_capi.ltypes([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)])) == [0, 1, 2, 3]


def test_synthetic_workloads_are_deterministic():
    a, b = synthetic.water_cluster(32), synthetic.water_cluster(32)
    assert a == b and len(a) == 96 and [z for z, _ in a[:3]] == [8, 1, 1]
    m = synthetic.perturbed_methanes(4)
    assert m[:2] == synthetic.perturbed_methanes(2) and len(m[0]) == 5


def _arrays(name):
    mol, basis, ne = systems.SYSTEMS[name]
    s = System_mol(mol, basis, ne)
    return _capi.SystemArrays(s.alpha, s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, False)


def test_work_decomposition_counts_every_quartet_once():
    """Tiles cover the packed pair-of-pairs space: every unique contracted quartet appears, diagonal tiles redundantly."""
    import bench
    for name in ("h2", "h2o_sto3g", "ch2o", "hcn"):
        a = _arrays(name)
        p1 = _capi.plan(a)
        assert p1["tiles"] == p1["tiles_total"] == p1["block_pairs"] * (p1["block_pairs"] + 1) // 2
        n = a.nbasis
        assert p1["quartets"] >= Tools.eri_vector_size(n)            # redundancy only inside diagonal tiles
        assert p1["quartets"] <= n ** 4
    s = bench.workload(32)
    a = _capi.SystemArrays(s.alpha, s.coef, s.xyz, s.l, s.charges, s.atom, s.list_contr, s.ne, False)
    p = _capi.plan(a)
    assert (p["blocks"], p["block_pairs"], p["tiles"]) == (56, 1596, 1274406)
    assert p["prim_quartets"] == p["quartets"] * 81
    assert bench.prim_quartets(s) == 25720140600                      # 2.572e10 unique primitive quartets


def test_rank_partition_is_a_disjoint_cover():
    a = _arrays("ch2o")
    whole = _capi.plan(a)
    for world in (2, 3, 8):
        parts = [_capi.plan(a, world, r) for r in range(world)]
        assert sum(p["tiles"] for p in parts) == whole["tiles"]
        assert sum(p["quartets"] for p in parts) == whole["quartets"]
        assert sum(p["prim_quartets"] for p in parts) == whole["prim_quartets"]
        assert max(p["tiles"] for p in parts) - min(p["tiles"] for p in parts) <= 16   # round-robin inside each class


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from diffiqult_b200 import _capi as capi, distributed
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    cb = distributed.make_allreduce(on_device=False)
    buf = np.arange(10, dtype=np.float64) * (rank + 1)
    cb(buf.ctypes.data, 10, None, None)                    # what the C library does with its partial J|K buffer
    a = _arrays("ch2o")
    mine = capi.plan(a, world, rank)
    cnt = np.array([mine["tiles"], mine["quartets"], mine["prim_quartets"]], dtype=np.float64)
    cb(cnt.ctypes.data, 3, None, None)
    q.put((rank, buf.tolist(), cnt.tolist(), cb.calls["n"]))
    dist.destroy_process_group()


def test_two_rank_allreduce_plumbing_over_gloo():
    import torch.multiprocessing as mp
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = _capi.plan(_arrays("ch2o"))
    for rank, buf, cnt, ncalls in out:
        assert buf == (np.arange(10) * 3.0).tolist()
        assert cnt == [whole["tiles"], whole["quartets"], whole["prim_quartets"]]
        assert ncalls == 2


def test_batch_host_logic():
    """diffiqult_b200.batch: structure test that gates the fused device path, and the contiguous rank slices (config 5)."""
    from diffiqult_b200 import batch
    ch4 = [System_mol(m, basis_set_3G_STO, 10, "ch4") for m in synthetic.perturbed_methanes(3)]
    assert batch.same_structure(ch4)
    h2o = System_mol(synthetic.H2O, basis_set_3G_STO, 10, "h2o")
    assert not batch.same_structure(ch4 + [h2o])
    cover = []
    for r in range(8):
        cover += list(batch.shard(1024, 8, r))
    assert cover == list(range(1024)) and len(batch.shard(1024, 8, 3)) == 128
    assert [len(batch.shard(10, 4, r)) for r in range(4)] == [3, 3, 3, 1]
    # without a GPU the fused entry point refuses like every other one (no CPU fallback)
    if _capi.lib().dq_device_count() == 0:
        with pytest.raises(_capi.DiffiqultB200Error, match="no CUDA device"):
            batch.rhf_grad_batch_fused(ch4, max_scf=5)


def test_bench_reference_arm_emits_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference algorithm on the host cores, a bounded sample): one JSON
    line with the contract's keys, no GPU needed."""
    import json
    import subprocess
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-waters", "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "eri_plus_gradient_primitive_quartets_per_s"
    assert line["value"] > 0 and line["unit"] == "primitive quartets/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("(H2O)_32") and line["n_gpus"] == 1 and line["steps"] == 1
