"""HyperpriDataset with the reference's constructor and item contract (reference src/dataset.py:25-298):
``{'image', 'mask', 'index', 'label'}`` per item, JSON-split or directory discovery, HSI band slice
``[hsi_lo:hsi_hi]``, RNG-state replay so image and mask get the same RandomCrop, the
``/255 if max > 10`` rule, mask binarisation.  The dataset stays CPU-only and fork/pickle-safe
(DataLoader workers, PLTrainer.py:470-471); the device-side half of ingest (slice/crop/flip/normalise ->
NHWC fp16) is ``hyperpri_b200.ops.hsi_ingest`` and is used when whole raw cubes are resident on the GPU.
"""
import json
import logging
import os
import pathlib

import numpy as np
import torch
from PIL import Image
from torch.utils.data import Dataset

from .. import envi


class HyperpriDataset(Dataset):
    N_BANDS = 299                                  # dataset.py:56

    def __init__(self, root, mode='RGB', img_transform=None, label_transform=None, subset: list = None,
                 label_subset: list = [3], unsqueeze_img=False, hsi_lo=0, hsi_hi=0, json_file: str = None,
                 json_verb=False, host_dtype=torch.float32):
        # host_dtype (extension, default = reference behaviour): torch.float16 makes HSI items half-precision cubes,
        # i.e. the fp32 -> fp16 rounding the device ingest would do happens before the PCIe copy (half the bytes,
        # bit-identical network input when no rescale / normalisation is configured).
        self.host_dtype = host_dtype
        self.class_list = list(subset) if subset is not None else ['Peanut', 'SweetCorn']
        ls = sorted(set(label_subset)) if label_subset else [0, 3]
        if 0 not in ls:
            ls.insert(0, 0)
        if ls[0] != 0:
            ls = [e - min(ls) for e in ls]
        self.label_subset = ls
        assert hsi_lo >= 0
        if hsi_hi <= 0:
            hsi_hi = self.N_BANDS + hsi_hi
        assert hsi_lo < hsi_hi
        self.root, self.mode = root, mode
        self.img_transform, self.label_transform = img_transform, label_transform
        self.unsqueeze_hsi = unsqueeze_img
        self.hsi_lo, self.hsi_hi = hsi_lo, hsi_hi
        self.files = []
        self.class_count = np.zeros(len(self.class_list), dtype=int)
        if not json_file:
            self._parse_train_dir()
        else:
            self.json_file = json_file
            self._parse_json_file(json_file, verbose=json_verb)
        # per-file sampling weights: under-represented classes are drawn more often (dataset.py:75-82)
        self.sample_weights = np.zeros(int(self.class_count.sum()))
        pos = 0
        for cnt in self.class_count:
            self.sample_weights[pos:pos + cnt] = 0 if cnt == 0 else self.class_count.max() / cnt
            pos += cnt

    # ---------------------------------------------------------------- discovery
    def _add(self, cls_idx, img, label, hdr=None, dat=None):
        item = {"img": img, "label": label if label.endswith('.png') else label.rsplit('.', 1)[0] + '.png'}
        if hdr is not None:
            item.update(hdr=hdr, dat=dat)
        self.files.append(item)
        self.class_count[cls_idx] += 1

    def _parse_train_dir(self):
        """Walk <root>/images/**; masks live under <root>/mask_files with '<name>_mask.png' (dataset.py:84-158)."""
        imgdir = os.path.join(self.root, 'images')
        for os_root, _dirs, files in os.walk(imgdir):
            cls = next((i for i, c in enumerate(self.class_list) if c in os_root), None)
            if cls is None or not files:
                continue
            parts = pathlib.Path(os_root).parts
            rel = parts[parts.index('images') + 1:]
            if self.mode.lower() == 'hsi':
                base = files[0].rsplit('.', 1)[0]
                hdr = os.path.join(os_root, "hinalea_hsi.hdr")
                rel_lbl = rel[:rel.index(base)] if base in rel else rel
                label = os.path.join(self.root, 'mask_files', *rel_lbl, f"{base}_mask.png")
                for pth in (hdr, os.path.join(os_root, base + ".dat"), label):
                    if not os.path.exists(pth):
                        raise FileNotFoundError(pth)
                self._add(cls, os.path.join(os_root, base + ".png"), label, hdr, os.path.join(os_root, base + ".dat"))
            else:
                for name in files:
                    label = os.path.join(self.root, 'mask_files', *rel, f"{name.rsplit('.', 1)[0]}_mask.png")
                    self._add(cls, os.path.join(os_root, name), label)

    def _parse_json_file(self, json_path, verbose=False):
        """Split files list rhizoboxes ("box*") with acquisition dates (dataset.py:160-244)."""
        with open(json_path, 'r') as f:
            d = json.load(f)
        for box, info in d.items():
            if not box.startswith("box") or not info['dates']:
                continue
            base_dir = f"{self.root}/{info['plant_folder']}_{info['resolution']}"
            cls = self.class_list.index(info['plant_folder'])
            for date in info['dates']:
                stem = f"{date}_{box}_ref"
                img = os.path.join(f"{base_dir}/{d['img_dir']}/", stem + ".png")
                label = os.path.join(f"{base_dir}/{d['mask_dir']}/", stem + "_mask.png")
                if self.mode.lower() == 'hsi':
                    hdr = os.path.join(f"{base_dir}/{d['hsi_dir']}/", "hinalea_hsi.hdr")
                    dat = os.path.join(f"{base_dir}/{d['hsi_dir']}/", stem + ".dat")
                    if not (os.path.exists(label) and os.path.exists(hdr) and os.path.exists(dat)):
                        if verbose:
                            logging.info(f"{stem}: missing HSI or mask file, skipped")
                        continue
                    self._add(cls, img, label, hdr, dat)
                else:
                    if not os.path.exists(img) or not os.path.exists(label):
                        if verbose:
                            print(f"Either {img} or {label} does not exist. Skipping...")
                        continue
                    self._add(cls, img, label)

    # ---------------------------------------------------------------- items
    def __len__(self):
        return len(self.files)

    def load_raw_cube(self, index) -> np.ndarray:
        """Whole cube, lines x samples x bands float32 (for device-side ingest)."""
        f = self.files[index]
        return envi.load(f['hdr'], f['dat'])

    def __getitem__(self, index):
        f = self.files[index]
        name = pathlib.PurePath(f["img"]).name.rsplit('.', 1)[0]
        mode = self.mode.lower()
        if mode == 'rgb':
            img = Image.open(f["img"]).convert('RGB')
        elif mode == 'gray':
            img = Image.open(f["img"]).convert('L').convert('RGB')
        else:
            cube = np.moveaxis(envi.load(f['hdr'], f['dat']), -1, 0)[self.hsi_lo:self.hsi_hi]
            if self.unsqueeze_hsi:
                cube = np.expand_dims(cube, 0)
            img = torch.tensor(cube)
        label = Image.open(f["label"]).convert("L")
        if mode != 'hsi' and img.size[0] < img.size[1]:
            img = img.transpose(method=Image.ROTATE_90)
            label = label.transpose(method=Image.ROTATE_90)
        state = torch.get_rng_state()              # the mask transform must draw the same crop (dataset.py:283-291)
        if self.img_transform is not None:
            img = self.img_transform(img)
            if img.max() > 10:
                img = img / 255
        torch.set_rng_state(state)
        if mode == 'hsi' and self.host_dtype != torch.float32:
            img = img.to(self.host_dtype)
        if self.label_transform is not None:
            label = self.label_transform(label)
        label = np.array(label) * 255
        label = np.where(label > 0, np.ones_like(label), np.zeros_like(label))
        return {'image': img, 'mask': label, 'index': name, 'label': f["label"]}
