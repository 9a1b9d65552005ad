"""End-to-end on a B200 through the reference's entry functions (PLTrainer.py:333-661): train_net on a tiny synthetic
ENVI dataset written to disk, then validate_net / test_net; the device validation sweep must give the same curve
as the host-side restatement of the reference's maths."""
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu

from hyperpri_b200 import envi, metrics as M                                   # noqa: E402
from hyperpri_b200.src import PLTrainer as T                                    # noqa: E402
from hyperpri_b200.src.Experiments.params_HyperPRI import ExpHyperspectralPRI   # noqa: E402


def _write_split(root, name, dates, h, w, seed):
    base = os.path.join(root, "Datasets", "HyperPRI", "Peanut_968x608")
    for d in ("rgb_files", "hsi_files", "mask_files"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    os.makedirs(os.path.join(root, "Datasets", "HyperPRI", "data_splits"), exist_ok=True)
    rs = np.random.RandomState(seed)
    for date in dates:
        stem = f"{date}_box37_ref"
        mask = np.zeros((h, w), np.uint8)
        mask[:, w // 3: w // 3 + 6] = 3                      # a vertical "root"
        cube = rs.random_sample((h, w, 299)).astype(np.float32) * 0.2
        cube[mask > 0, 100:180] += 0.6                       # roots are bright in the middle bands: learnable
        envi.save(os.path.join(base, "hsi_files", "hinalea_hsi.hdr"), os.path.join(base, "hsi_files", stem + ".dat"), cube)
        Image.fromarray(mask).save(os.path.join(base, "mask_files", stem + "_mask.png"))
        Image.fromarray((rs.random_sample((h, w, 3)) * 255).astype(np.uint8)).save(os.path.join(base, "rgb_files", stem + ".png"))
    js = {"img_dir": "rgb_files", "hsi_dir": "hsi_files", "mask_dir": "mask_files", "notes": "synthetic",
          "box37": {"plant_folder": "Peanut", "resolution": "968x608", "box_no": 37, "phenotype": 1, "dates": dates,
                    "weights": None}}
    with open(os.path.join(root, "Datasets", "HyperPRI", "data_splits", f"{name}1.json"), "w") as f:
        json.dump(js, f)


def test_train_validate_test_roundtrip(tmp_path):
    root = str(tmp_path)
    _write_split(root, "train", ["20220701", "20220702", "20220703", "20220704"], 32, 48, 0)
    _write_split(root, "val", ["20220711", "20220712"], 32, 48, 1)
    torch.manual_seed(0)
    p = ExpHyperspectralPRI(root, split_no=1, seed_num=0, comet_logging=False)
    p.epochs, p.patch_size = 6, (32, 48)
    trainer = T.train_net(p)
    assert len(trainer.history) == 6 and all("tr_loss" in h and "val_loss" in h and "tr_dice" in h for h in trainer.history)
    hist = [h["tr_loss"] for h in trainer.history]
    assert trainer.model.m_network.first_conv.weight.is_cuda
    assert hist[-1] < hist[0]
    opt_probe = trainer.model.configure_optimizers()              # FusedAdam wired to the engine's overflow flag
    assert getattr(opt_probe, "found_inf", None) is not None and opt_probe.skipped_steps() == 0
    # device sweep (default on CUDA) vs the host path the reference takes (concatenate predictions, torch ops)
    prec_d, rec_d, thr_d = T.validate_net(p.get_val_data(), p, pl_trainer=trainer)
    best = float(trainer.model.threshold)
    # the reference evaluates the best val_loss checkpoint, not the trainer's last-epoch weights (PLTrainer.py:476)
    best_model = T.load_val_model(p)
    ck = torch.load(os.path.join(p.save_path, "Checkpoints", "best.ckpt"), weights_only=False)
    assert ck["epoch"] == min(trainer.history, key=lambda h: h["val_loss"])["epoch"]
    logits, masks = T._collect(best_model, torch.utils.data.DataLoader(p.get_val_data(), batch_size=2), trainer)
    prec_h, rec_h, thr_h = M.binned_pr_curve(torch.sigmoid(logits), masks, 500)
    if prec_h[-2] < 1e-6:
        prec_h[-2] = (1 + prec_h[-3]) / 2
    assert torch.equal(thr_d.cpu(), thr_h) and torch.allclose(prec_d.cpu(), prec_h, atol=1e-6)
    assert torch.allclose(rec_d.cpu(), rec_h, atol=1e-6)
    assert 0.0 <= best <= 1.0 and abs(best * 100 - round(best * 100)) < 1e-4
    out = T.test_net(p.get_test_data(), p, best, pl_trainer=trainer)
    assert set(out) == {"acc", "dice", "pos_iou", "avg_prec"} and all(0.0 <= v <= 1.0 for v in out.values())
    assert out["dice"] > 0.5                                 # six epochs on a trivially separable signal


def test_deterministic_knob_makes_train_net_bit_reproducible(tmp_path):
    """params.deterministic = True (the reference's Trainer(deterministic='warn'), PLTrainer.py:430,439,447): two
    train_net runs from the same seed end with identical loss histories and identical weights."""
    from hyperpri_b200 import ops
    runs = []
    try:
        for r in range(2):
            root = str(tmp_path / f"run{r}")
            _write_split(root, "train", ["20220701", "20220702", "20220703", "20220704"], 32, 48, 0)
            _write_split(root, "val", ["20220711", "20220712"], 32, 48, 1)
            torch.manual_seed(0)
            p = ExpHyperspectralPRI(root, split_no=1, seed_num=0, comet_logging=False)
            p.epochs, p.patch_size, p.deterministic = 3, (32, 48), True
            trainer = T.train_net(p)
            assert ops.DETERMINISTIC and ops.DETERMINISTIC_BWD
            runs.append(([(h["tr_loss"], h["val_loss"]) for h in trainer.history],
                         {k: v.detach().clone() for k, v in trainer.model.m_network.state_dict().items()}))
    finally:
        ops.set_deterministic(False)
    assert runs[0][0] == runs[1][0]
    for k, v in runs[0][1].items():
        assert torch.equal(v, runs[1][1][k]), k


@pytest.mark.gpu
def test_device_prefetcher_order_contents_and_slot_reuse():
    """Batches arrive on the device in order, bit-identical, one copy ahead; the two slots are recycled without the
    consumer ever seeing a later batch's bytes (each batch is reduced on the device well after the next copy was queued)."""
    import torch
    from hyperpri_b200.prefetch import DevicePrefetcher
    g = torch.Generator().manual_seed(3)
    batches = [{"image": torch.rand((2, 8, 64, 96), generator=g).half() if i % 2 else torch.rand((2, 8, 64, 96), generator=g),
                "mask": (torch.rand((2, 1, 64, 96), generator=g) > 0.5).float(), "index": torch.tensor([i]), "label": str(i)}
               for i in range(7)]
    pf = DevicePrefetcher(batches, "cuda")
    sums, seen = [], []
    for b in pf:
        assert b["image"].is_cuda and b["mask"].is_cuda and not b["index"].is_cuda and isinstance(b["label"], str)
        torch.cuda._sleep(2_000_000)                       # keep the consumer busy while the next copy lands
        sums.append((b["image"].clone(), b["mask"].clone()))   # stream-ordered after the sleep
        seen.append(int(b["index"]))
    torch.cuda.synchronize()
    assert seen == list(range(7))
    for (im, mk), b in zip(sums, batches):
        assert im.dtype == b["image"].dtype and torch.equal(im.cpu(), b["image"]) and torch.equal(mk.cpu(), b["mask"])
    assert pf.h2d_bytes == sum(b["image"].numel() * b["image"].element_size() + b["mask"].numel() * 4 for b in batches)
