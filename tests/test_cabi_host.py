"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares (no compute calls -- there is no GPU here), the ctypes prototypes agree with the
header, the module reproduces the reference's state_dict keys and seeded initial weights, and
the product path refuses to run without a CUDA device instead of falling back."""
import argparse
import ctypes
import os
import re

import pytest
import torch

from helpers import GOLDEN_NAMES, ROOT, Golden


def _header_decls():
    text = open(os.path.join(ROOT, "include", "c2dsr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(c2dsr_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        decls[name] = n
    return decls


@pytest.fixture(scope="module")
def built_lib():
    from c2dsr_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    from c2dsr_b200 import _cabi
    decls = _header_decls()
    assert len(decls) >= 30
    lib = ctypes.CDLL(built_lib)
    for name, n_params in decls.items():
        assert hasattr(lib, name), f"{name} declared in include/c2dsr_b200.h but not exported"
        assert name in _cabi._PROTOS, f"{name} has no ctypes prototype"
        assert len(_cabi._PROTOS[name][1]) == n_params, f"{name}: ctypes prototype has the wrong arity"
    assert set(_cabi.EXPORTS) == set(decls)
    assert lib.c2dsr_abi_version() == 1


def test_size_queries_need_no_device(built_lib):
    from c2dsr_b200 import _cabi
    assert _cabi.query("c2dsr_score_ldz", 29207) == 29208
    assert _cabi.query("c2dsr_gather_bwd_workspace_bytes", 3840, 256, 64094, 15) > 2 * 64094 * 4
    assert _cabi.query("c2dsr_encoder_saved_floats", 3840, 256, 1, 1) == 3840 * (9 * 256 + 1 + 4) + 3840 * 258


def test_no_cpu_fallback(built_lib):
    """Compute entry points must raise without a CUDA device; nothing silently runs on the host."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from c2dsr_b200 import _cabi, ops
    with pytest.raises(_cabi.C2dsrError):
        _cabi.call("c2dsr_axpby", None, None, None, 0, 1.0, 0.0, None)
    with pytest.raises(Exception):
        ops.score_shard(torch.zeros(2, 4), torch.zeros(3, 4), torch.zeros(3))


def _args_from(hp):
    return argparse.Namespace(**hp, device=torch.device("cpu"))


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_module_keys_and_seeded_init_equal_reference(name):
    """Same containers built in the same order => same state_dict keys AND the reference's own
    initial weights for the same torch seed (models/C2DSR.py:9-57)."""
    from c2dsr_b200.c2dsr import C2DSR
    g = Golden(name)
    torch.manual_seed(g.hp["seed"])
    model = C2DSR(_args_from(g.hp), g.adj("share"), g.adj("spec"))
    ref = g.group("init")
    sd = model.state_dict()
    assert set(sd) == set(ref)
    for k, v in ref.items():
        assert torch.equal(sd[k], v), k
    # dead prototype layers exist, and live layers are separate tensors (Q3)
    assert "attn_a.encoder_layer.linear1.weight" in sd
    assert model.attn_a.encoder_layer.linear1.weight is not model.attn_a.encoder.layers[0].linear1.weight
    model.load_state_dict(g.group("final"))          # a reference checkpoint loads unchanged


def test_main_flags_cover_reference():
    from c2dsr_b200.main import FLAGS
    ref_flags = {"data", "len_rec", "use_raw", "n_neg_sample", "zip_ee", "d_latent", "disable_embed_l2",
                 "shared_item_embed", "d_bias", "n_gnn", "dropout_gnn", "n_attn", "n_head", "dropout_attn",
                 "norm_first", "lr", "lr_decay", "l2", "lr_gamma", "lr_step", "n_lr_decay", "decay_epoch",
                 "max_grad_norm", "len_max", "lambda_loss", "cuda", "seed", "n_epoch", "batch_size",
                 "batch_size_eval", "num_workers", "es_patience"}
    assert ref_flags <= {n.lstrip("-") for n, _ in FLAGS}
    defaults = {n.lstrip("-"): kw.get("default") for n, kw in FLAGS}
    assert (defaults["d_latent"], defaults["batch_size"], defaults["seed"], defaults["l2"]) == (128, 512, 3407, 5e-4)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the unmodified reference from oracle/_ref when it is there -- placed by
    oracle/make_ref.py / __graft_entry__.build() -- else the oracle port) prints the driver's JSON line: same metric /
    unit / config keys as the GPU arm, impl = reference, exactly the requested steps, e2e without transfers."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    z = json.loads(r.stdout.strip().splitlines()[-1])
    assert z["impl"] == "reference" and z["metric"] == "train_seqs_per_sec" and z["unit"] == "seq/s"
    assert z["value"] > 0 and z["higher_is_better"] is True and z["n_gpus"] == 1 and z["steps"] == 3
    assert z["e2e"] == {"value": z["value"], "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = z["cpu_baseline"]
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "trainer.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == z["value"]
    assert "3 timed steps" in cb["sample"]
    # the same workload description as the GPU arm prints (bench.common_config)
    import argparse
    import bench
    import torch
    hp = bench.hyper(bench.WORKLOADS["tiny"], 0.2, torch.device("cpu"))
    assert z["config"] == bench.common_config(hp, 0.2, 1)


def test_oracle_port_equals_reference_when_present():
    """oracle/_ref (the unmodified reference, git-ignored) against the oracle port on one seeded tiny step: the port
    is what the parity tests use, the reference is what the CPU arm times -- they must be the same algorithm."""
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "trainer.py")):
        pytest.skip("oracle/_ref not built (run python oracle/make_ref.py where /root/reference exists)")
    import tempfile
    import numpy as np
    import torch
    import bench
    import ref_run
    import c2dsr_oracle as oracle
    hp = bench.hyper(bench.WORKLOADS["tiny"], 0.0, torch.device("cpu"))
    adj, fields, ev = bench.make_workload(hp, 2, 1, seed=0)
    root = tempfile.mkdtemp(prefix="c2dsr_ref_test_")
    ref_run.write_processed(root, hp.dataset, hp.n_item_a, hp.n_item_b, fields, ev, adj)
    h = {k: v for k, v in vars(hp).items() if isinstance(v, (int, float, bool, str))}
    tr, args = ref_run.load_trainer(root, h)
    otr = oracle.OracleTrainer({k: v.detach().clone() for k, v in tr.model.state_dict().items()},
                               adj[0].coalesce(), adj[1].coalesce(), h)
    B = hp.batch_size
    batch = tuple(torch.from_numpy(np.ascontiguousarray(fields[:B, f])) for f in range(14))
    tr.model.train(); tr.optimizer.zero_grad(); otr.zero_grad()
    tr.model.convolve_graph()
    ref = [float(x) for x in tr.train_batch(batch)]
    got = [float(x) for x in otr.train_batch(batch, training=True)]
    np.testing.assert_allclose(got, ref, rtol=2e-6)


def test_header_is_plain_c():
    """include/c2dsr_b200.h is the drop-in boundary: it must compile as C99 (no C++ or torch types) and as C++."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "c2dsr_b200.h")
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_struct_layouts_match_the_ctypes_mirror(tmp_path):
    """The structs that cross the boundary by value or as device tables (c2dsr_adam_tensor, c2dsr_peer_tensor,
    c2dsr_peer_map) have the same size and field offsets in C as in c2dsr_b200/_cabi.py."""
    import ctypes
    import shutil
    import subprocess
    from c2dsr_b200 import _cabi
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "layout.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "c2dsr_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu\\n", sizeof(c2dsr_adam_tensor), offsetof(c2dsr_adam_tensor, acc),
           offsetof(c2dsr_adam_tensor, vmax), offsetof(c2dsr_adam_tensor, n));
    printf("%zu %zu\\n", sizeof(c2dsr_peer_tensor), offsetof(c2dsr_peer_tensor, offset));
    printf("%zu %zu %zu %zu %zu %d\\n", sizeof(c2dsr_peer_map), offsetof(c2dsr_peer_map, grad),
           offsetof(c2dsr_peer_map, param), offsetof(c2dsr_peer_map, grad_mc), offsetof(c2dsr_peer_map, param_mc),
           C2DSR_MAX_PEERS);
    return 0;
}
''')
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [[int(x) for x in line.split()] for line in subprocess.run([str(exe)], capture_output=True,
                                                                      text=True).stdout.strip().splitlines()]
    A, P, M = _cabi.AdamTensor, _cabi.PeerTensor, _cabi.PeerMap
    assert got[0] == [ctypes.sizeof(A), A.acc.offset, A.vmax.offset, A.n.offset]
    assert got[1] == [ctypes.sizeof(P), P.offset.offset]
    assert got[2] == [ctypes.sizeof(M), M.grad.offset, M.param.offset, M.grad_mc.offset, M.param_mc.offset,
                      _cabi.MAX_PEERS]
