"""CPU: host-side logic of the drop-in -- the C-ABI library loads and exports every declared symbol, state-dict
schemas equal the reference's, parameter classes keep names/defaults, dataset + ENVI reader + metrics."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch
from PIL import Image

import hyperpri_oracle as O
from hyperpri_b200 import _lib, envi, metrics as M
from hyperpri_b200.src.Experiments import models as mdl
from hyperpri_b200.src.Experiments.params_HyperPRI import ExpHyperspectralPRI, ExpRedGreenBluePRI
from hyperpri_b200.src.dataset import HyperpriDataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    import __graft_entry__ as ge
    ge.build()
    hdr = open(os.path.join(ROOT, "include", "hyperpri_b200.h")).read()
    declared = set(re.findall(r"\b(hpri_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.hpri_abi_version() == 7
    assert ctypes.sizeof(_lib.View) == 56          # ptr, 4 ints, 3 int64, dtype (+pad)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhyperpri_b200.so")
    with pytest.raises(_lib.NativeLibraryMissing):
        _lib.lib()


@pytest.mark.parametrize("net,schema,nparam", [
    (lambda: mdl.UNet(3, 1, bilinear=False), lambda: O.unet_schema(3, 1, "unet"), 31043521),
    (lambda: mdl.CubeNET(238, 1, bilinear=False), lambda: O.unet_schema(1, 1, "cube", 238), 31178881),
    (lambda: mdl.SpectralUNET(238, 1, bn_feats=1650), lambda: O.spectral_schema(238, 1, 1650), 30388051),
    (lambda: mdl.SpectralUNET(238, 1, bn_feats=32, bnorm=False), lambda: O.spectral_schema(238, 1, 32, bnorm=False),
     32 * 238 + 32 + 5 * (32 * 32 + 32) + 3 * (64 * 32 + 32) + 65),
])
def test_state_dict_schema_matches_reference(net, schema, nparam):
    n, s = net(), schema()
    sd = n.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v) for k, v in s.items()}
    assert sum(p.numel() for p in n.parameters()) == nparam      # SURVEY.md section 4 known answers
    n.load_state_dict(O.synth_state_dict(s, 0))                    # reference-style state dict loads


def test_cubenet_aliases_first_conv():
    n = mdl.CubeNET(238, 1, bilinear=False)
    assert n.inc[0] is n.first_conv and n.n_channels == 1 and n.depth == 238
    assert len(n.state_dict()) == 138 and len(list(n.named_parameters())) == 82


def test_factories_and_unsupported_flags():
    p = dict(channels=3, bilinear=False, feature_extraction=False, use_attention=False, hsi_lo=25, hsi_hi=263,
             spectral_bn_size=1650, **{"3d_featmaps": 64})
    assert isinstance(mdl.initialize_model("UNET", 1, p), mdl.UNet)
    assert isinstance(mdl.initialize_model("CubeNET", 1, p), mdl.CubeNET)
    assert mdl.initialize_model("SpectralUNET", 1, p).layer_feats == [1650] * 5
    with pytest.raises(RuntimeError, match="Invalid model"):
        mdl.initialize_model("nope", 1, p)
    assert mdl.translate_load_dir("SpectralUNET", p) == "SpectralUNET_1650"
    assert mdl.translate_load_dir("CubeNET", p) == "CubeNET_64" and mdl.translate_load_dir("UNET", p) == "UNET"
    with pytest.raises(ValueError):                # does not run in the reference either (models.py:195-196)
        mdl.CubeNET(238, 1, first_depth=32, bilinear=True)
    bsd = mdl.UNet(3, 1).state_dict()               # the reference's default: bilinear=True
    assert {k: tuple(v.shape) for k, v in bsd.items()} == {k: tuple(v) for k, v in O.unet_schema(3, 1, "unet", bilinear=True).items()}
    assert bsd["down4.maxpool_conv.1.double_conv.3.weight"].shape == (512, 512, 3, 3) and "up1.up.weight" not in bsd
    # first_depth != 64 and use_attention are built: same state-dict schema as the reference (models.py:193-199)
    net = mdl.CubeNET(238, 1, first_depth=32, bilinear=False, use_attention=True)
    sd = net.state_dict()
    assert sd["first_conv.weight"].shape == (32, 1, 238, 3, 3) and sd["upsample4.weight"].shape == (128, 64, 2, 2)
    assert sd["upconv4.double_conv.0.weight"].shape == (64, 96, 3, 3) and "up4.up.weight" not in sd
    assert sd["up1.conv.double_conv.0.weight"].shape == (512, 512, 3, 3)          # attention: skip * up, Cin / 2
    assert mdl.translate_load_dir("CubeNET", dict(p, **{"3d_featmaps": 32})) == "CubeNET_32"


def test_cpu_forward_is_refused_not_emulated():
    n = mdl.UNet(3, 1, bilinear=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        n(torch.zeros(1, 3, 16, 16))


def test_param_classes_keep_reference_knobs(tmp_path):
    p = ExpHyperspectralPRI(str(tmp_path), split_no=3, seed_num=1, comet_logging=False)
    assert (p.hsi_lo, p.hsi_hi, p.channels, p.spectral_bn_size, p.cube_featmaps) == (25, 263, 238, 1650, 64)
    assert p.b_size == {'train': 2, 'val': 2, 'test': 2} and p.patch_size == (608, 968) and p.run_num == 13
    assert p.model_name == "CubeNET" and p.model_param_str == "CubeNET_64" and p.optimizer == "adam"
    assert p.learn_rate == 0.001 and p.threshold == 0.5 and isinstance(p.criterion, torch.nn.BCEWithLogitsLoss)
    assert p.save_path.endswith("Saved_Models/HSI/CubeNET_64/Run_13/")
    p.change_network_param("SpectralUNET", str(tmp_path), 3, model_params={"spectral_bn_size": 64, "nonexistent": 1})
    assert p.model_param_str == "SpectralUNET_64" and not hasattr(p, "nonexistent")
    assert isinstance(p.get_network(), mdl.SpectralUNET)
    r = ExpRedGreenBluePRI(str(tmp_path), comet_logging=False)
    assert r.channels == 3 and r.b_size['test'] == 1 and isinstance(r.get_network(), mdl.UNet)


def _make_dataset(root, n_dates=2, h=12, w=20):
    base = os.path.join(root, "Peanut_968x608")
    for d in ("rgb_files", "hsi_files", "mask_files", "../data_splits"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    dates = [f"2022070{i}" for i in range(n_dates)]
    rs = np.random.RandomState(0)
    cubes = {}
    for date in dates:
        stem = f"{date}_box37_ref"
        cube = rs.random_sample((h, w, 299)).astype(np.float32)
        cubes[stem] = cube
        envi.save(os.path.join(base, "hsi_files", "hinalea_hsi.hdr"), os.path.join(base, "hsi_files", stem + ".dat"), cube)
        Image.fromarray((rs.random_sample((h, w)) > 0.8).astype(np.uint8) * 3).save(os.path.join(base, "mask_files", stem + "_mask.png"))
        Image.fromarray((rs.random_sample((h, w, 3)) * 255).astype(np.uint8)).save(os.path.join(base, "rgb_files", stem + ".png"))
    js = {"img_dir": "rgb_files", "hsi_dir": "hsi_files", "mask_dir": "mask_files", "notes": "x",
          "box37": {"plant_folder": "Peanut", "resolution": "968x608", "box_no": 37, "phenotype": 1, "dates": dates + ["19990101"],
                    "weights": None}}
    jp = os.path.join(root, "data_splits", "val1.json")
    with open(jp, "w") as f:
        json.dump(js, f)
    return jp, cubes


@pytest.mark.parametrize("inter", ["bil", "bsq", "bip"])
def test_envi_roundtrip(tmp_path, inter):
    cube = np.random.RandomState(1).random_sample((5, 7, 9)).astype(np.float32)
    envi.save(str(tmp_path / "a.hdr"), str(tmp_path / "a.dat"), cube, inter)
    assert np.array_equal(envi.load(str(tmp_path / "a.hdr"), str(tmp_path / "a.dat")), cube)


def test_dataset_hsi_items_match_oracle_ingest(tmp_path):
    from torchvision import transforms
    jp, cubes = _make_dataset(str(tmp_path))
    ds = HyperpriDataset(root=str(tmp_path), mode='HSI', img_transform=None,
                         label_transform=transforms.Compose([transforms.ToTensor()]), unsqueeze_img=True, hsi_lo=25,
                         hsi_hi=263, json_file=jp)
    assert len(ds) == 2                         # the date without files is skipped
    it = ds[0]
    assert set(it) == {'image', 'mask', 'index', 'label'} and it['index'] == "20220700_box37_ref"
    assert tuple(it['image'].shape) == (1, 238, 12, 20) and it['image'].dtype == torch.float32
    assert np.array_equal(it['image'].numpy(), O.ingest_hsi(cubes[it['index']], 25, 263, True))
    m = np.asarray(it['mask'])
    assert m.shape == (1, 12, 20) and set(np.unique(m)) <= {0.0, 1.0}
    # crop transform: image and mask share the drawn window (RNG-state replay)
    ds2 = HyperpriDataset(root=str(tmp_path), mode='HSI', img_transform=transforms.Compose([transforms.RandomCrop((8, 10))]),
                          label_transform=transforms.Compose([transforms.RandomCrop((8, 10)), transforms.ToTensor()]),
                          unsqueeze_img=False, hsi_lo=25, hsi_hi=263, json_file=jp)
    torch.manual_seed(5)
    it2 = ds2[1]
    full = O.ingest_hsi(cubes[it2['index']], 25, 263, False)
    found = [(i, j) for i in range(5) for j in range(11) if np.array_equal(full[:, i:i + 8, j:j + 10], it2['image'].numpy())]
    assert len(found) == 1
    i, j = found[0]
    full_mask = np.asarray(ds[1]['mask'])
    assert np.array_equal(np.asarray(it2['mask']), full_mask[:, i:i + 8, j:j + 10])


def test_metrics_against_bruteforce():
    g = torch.Generator().manual_seed(0)
    probs = torch.rand(5000, generator=g)
    tgt = (torch.rand(5000, generator=g) < probs * 0.6).long()
    prec, rec, thr = M.binned_pr_curve(probs, tgt, 500)
    assert prec.shape == (501,) and rec.shape == (501,) and thr.shape == (500,)
    for i in (0, 17, 250, 499):
        pred = probs >= thr[i]
        tp = (pred & (tgt > 0)).sum().item(); fp = (pred & (tgt == 0)).sum().item()
        assert abs(prec[i].item() - (tp / (tp + fp) if tp + fp else 0.0)) < 1e-6
        assert abs(rec[i].item() - tp / (tgt > 0).sum().item()) < 1e-6
    assert prec[-1] == 1 and rec[-1] == 0 and 0 < M.average_precision(prec, rec) < 1
    c = M.confusion_counts(probs > 0.5, tgt)
    tp, fp, fn, tn = [v.item() for v in c]
    assert tp + fp + fn + tn == 5000
    assert abs(M.dice(*c).item() - 2 * tp / (2 * tp + fp + fn)) < 1e-6
    assert abs(M.jaccard(*c).item() - tp / (tp + fp + fn)) < 1e-6
    assert O.seg_counts(torch.logit(probs), tgt.float()) == (tp, fp, fn, tn)


def test_device_prefetcher_is_a_passthrough_on_cpu():
    import torch
    from hyperpri_b200.prefetch import DevicePrefetcher
    batches = [{"image": torch.full((1, 3, 4, 4), float(i)), "mask": torch.zeros(1, 1, 4, 4), "index": i} for i in range(3)]
    out = list(DevicePrefetcher(batches, "cpu"))
    assert len(out) == 3 and all(o is b for o, b in zip(out, batches))
    assert len(DevicePrefetcher(batches, "cpu")) == 3


def _write_zero2_dir(root, sd, buffer_names, world, groups=2, tag="checkpoint"):
    """A directory in the DeepSpeed ZeRO-2 layout hyperpri_b200/zero_ckpt.py documents (synthetic: no DeepSpeed here)."""
    from collections import OrderedDict
    d = os.path.join(root, tag)
    os.makedirs(d)
    with open(os.path.join(root, "latest"), "w") as f:
        f.write(tag)
    params = [(k, v) for k, v in sd.items() if k not in buffer_names]
    cut = len(params) // groups
    grp = [params[:cut], params[cut:]] if groups == 2 else [params]
    shapes = [OrderedDict((k, v.shape) for k, v in g) for g in grp]
    torch.save({"module": {k: (v.half() if v.dtype.is_floating_point else v) for k, v in sd.items()},
                "buffer_names": list(buffer_names), "param_shapes": shapes, "shared_params": {}, "ds_version": "0.14.2"},
               os.path.join(d, "mp_rank_00_model_states.pt"))
    parts = [[] for _ in range(world)]
    for g in grp:
        flat = torch.cat([v.reshape(-1).float() for _, v in g])
        pad = (-flat.numel()) % (2 * world)
        flat = torch.cat([flat, torch.zeros(pad)])
        for r, chunk in enumerate(flat.chunk(world)):
            parts[r].append(chunk.clone())
    for r in range(world):
        torch.save({"optimizer_state_dict": {"zero_stage": 2, "partition_count": world,
                                             "single_partition_of_fp32_groups": parts[r]}},
                   os.path.join(d, f"bf16_zero_pp_rank_{r}_mp_rank_00_optim_states.pt"))


@pytest.mark.parametrize("world", [1, 2, 3])
def test_zero2_checkpoint_directory_import(tmp_path, world):
    """consolidate_deepspeed_two (PLTrainer.py:186-216) on a directory written to the documented ZeRO-2 layout: fp32
    masters come back exactly (not the 16-bit "module" copies), buffers are kept, keys lose the Lightning prefix."""
    from hyperpri_b200.zero_ckpt import consolidate_deepspeed_two, fp32_state_dict_from_zero2
    net = mdl.SpectralUNET(238, 1, bn_feats=24)
    sd = {k: (torch.randn_like(v) if v.dtype.is_floating_point else v + 3) for k, v in net.state_dict().items()}
    pref = {"_forward_module.m_network." + k: v for k, v in sd.items()}
    buffers = [k for k in pref if "running_" in k or "num_batches" in k]
    _write_zero2_dir(str(tmp_path), pref, buffers, world)
    got = consolidate_deepspeed_two(str(tmp_path))
    assert set(got) == set(sd)
    for k, v in sd.items():
        if "running_" in k:
            assert torch.equal(got[k], v.half().float()), k        # buffers live in the 16-bit module copy
        else:
            assert torch.equal(got[k], v), k
    net.load_state_dict(got)
    full = fp32_state_dict_from_zero2(str(tmp_path), tag="checkpoint")
    assert all(k.startswith("_forward_module.m_network.") for k in full)
    os.remove(os.path.join(str(tmp_path), "checkpoint", f"bf16_zero_pp_rank_{world - 1}_mp_rank_00_optim_states.pt"))
    with pytest.raises((ValueError, FileNotFoundError)):
        consolidate_deepspeed_two(str(tmp_path))


@pytest.mark.parametrize("kind,kw", [
    ("UNET", {}), ("UNET", {"use_attention": True}), ("UNET", {"bilinear": True}), ("UNET", {"bilinear": True, "use_attention": True}),
    ("CubeNET", {}), ("CubeNET", {"first_depth": 16}), ("CubeNET", {"first_depth": 128, "use_attention": True}),
    ("CubeNET", {"bilinear": True}), ("CubeNET", {"use_attention": True}),
])
def test_engine_channel_plan_matches_parameters(kind, kw):
    """Host-side plan of UNetEngine for every constructor-flag combination (no kernels run: the engine is built on the
    CPU device): each conv unit's (cin, cout) equals its parameter's shape, the flat gradient arena holds every
    parameter exactly once and the nine buckets tile it in backward-completion order."""
    from hyperpri_b200 import engine as E
    kw = dict({"bilinear": False}, **kw)
    net = mdl.UNet(3, 1, **kw) if kind == "UNET" else mdl.CubeNET(238, 1, **kw)
    table = net._tensor_table()
    eng = E.UNetEngine(table, "unet" if kind == "UNET" else "cube", 3 if kind == "UNET" else 238, torch.device("cpu"),
                       attention=kw.get("use_attention", False), first_depth=kw.get("first_depth", 64),
                       bilinear=kw["bilinear"])
    assert eng._side is None                                           # no side stream without CUDA
    units = [L for grp in list(eng.enc) + [eng.dec[l] for l in range(4)] for L in grp]
    assert len(units) == 18
    for L in units:
        w = table[L.conv + ".weight"]
        cin_true = getattr(L, "true_cin", L.cin)
        got = (w.shape[0], w.shape[1] * (w.shape[2] if w.dim() == 5 else 1))
        assert got == (L.cout, cin_true), (L.conv, tuple(w.shape), L.cin, L.cout)
        assert table[L.bn + ".weight"].shape == (L.cout,)
    if kw["bilinear"]:
        assert not eng.up and eng.CE[4] == 512 and eng.D == [64, 64, 128, 256]
    else:
        for l, up in eng.up.items():
            w = table[eng.up_name[l] + ".weight"]
            assert tuple(w.shape) == (up.spec.cin, up.spec.cout, 2, 2), (l, tuple(w.shape))
    names = {k for k, p in net.named_parameters()}
    arena_names = set(eng.grads)
    alias = {"inc.0.weight", "inc.0.bias"} if kind == "CubeNET" else set()
    assert arena_names | alias >= names and arena_names <= names | alias
    assert eng.arena.numel() == sum(p.numel() for p in net.parameters())
    b = eng.bucket_bounds
    assert len(b) == 9 and b[0][0] == 0 and b[-1][1] == eng.arena.numel()
    assert all(b[i][1] == b[i + 1][0] and b[i][1] > b[i][0] for i in range(8))


def test_ctypes_structs_match_the_c_header_layout(tmp_path):
    """The ctypes mirrors in hyperpri_b200/_lib.py against the structs of include/hyperpri_b200.h as gcc lays them out:
    size and the offset of every field (an ABI drift between the header and the Python binding would corrupt launches
    silently)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = {"hpri_view_t": (_lib.View, ["ptr", "n", "h", "w", "c", "pix_stride", "row_stride", "img_stride", "dtype"]),
               "hpri_bn_fin_t": (_lib.BnFin, ["gamma", "beta", "conv_bias", "running_mean", "running_var", "num_batches_tracked",
                                              "scale", "shift", "save_mean", "save_invstd", "counter", "count", "momentum",
                                              "eps", "partials"]),
               "hpri_bn_bwd_t": (_lib.BnBwd, ["x", "scale", "shift", "save_mean", "save_invstd", "sums"]),
               "hpri_conv3x3_job_t": (_lib.Conv3x3Job, ["w", "dst_fwd", "dst_dgrad", "grad_packed", "grad_dst", "cout", "cin",
                                                        "fwd_dtype", "dgrad_dtype", "tile0", "kind"]),
               "hpri_adam_job_t": (_lib.AdamJob, ["param", "grad", "exp_avg", "exp_avg_sq", "numel", "block0"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hyperpri_b200.h"', 'int main(void) {']
    for cname, (_, fields) in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    want = {}
    for line in out.splitlines():
        cname, key, val = line.split()
        want[(cname, key)] = int(val)
    for cname, (cls, fields) in structs.items():
        assert ctypes.sizeof(cls) == want[(cname, "size")], cname
        for f in fields:
            assert getattr(cls, f).offset == want[(cname, f)], (cname, f)


def test_deterministic_switch_levels():
    """ops.set_deterministic: off / forward statistics only / forward + backward, mirrored into the library's own switch
    (hpri_set_deterministic accepts the call without a GPU) and read by the engine's launch plan."""
    from hyperpri_b200 import engine as E, ops
    try:
        ops.set_deterministic(True, backward=False)
        assert ops.DETERMINISTIC and not ops.DETERMINISTIC_BWD and E._EngineBase._wgrad_splits() == 0
        ops.set_deterministic(True)
        assert ops.DETERMINISTIC and ops.DETERMINISTIC_BWD and E._EngineBase._wgrad_splits() == 1
        net = mdl.UNet(3, 1, bilinear=False)
        eng = E.UNetEngine(net._tensor_table(), "unet", 3, torch.device("cpu"))
        eng.ws = {"H": [608, 304, 152, 76, 38], "W": [968, 484, 242, 121, 60]}
        assert not any(eng._fusable(l) for l in range(4))          # no fused dgrad + BatchNorm-backward epilogue
        ops.set_deterministic(False, backward=True)
        assert not ops.DETERMINISTIC and not ops.DETERMINISTIC_BWD and E._EngineBase._wgrad_splits() == 0
        assert all(eng._fusable(l) for l in range(4))              # the production plan at BASELINE size
    finally:
        ops.set_deterministic(False)
