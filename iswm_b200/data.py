"""Host -> device input staging for the hot path.

The reference loop copies every batch synchronously right before the forward pass
(`images.to(device, float32); labels.to(device, long)`, train.py:1039-1040), so the GPU idles for the
whole PCIe transfer (84 MB per step at batch 16, 512x512). `HostBatchPrefetcher` issues the copy of batch
i+1 on a side stream while step i computes; the consumer's stream waits only on the copy it is about to use.
Works with any iterable of `(images, labels)` tuples or `{'image'|'img': ..., 'mask': ...}` dicts of host
tensors (pinned memory makes the copies asynchronous; pageable tensors still work, synchronously).
"""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import torch


class HostBatchPrefetcher:
    def __init__(self, loader: Iterable, device, image_dtype=torch.float32, label_dtype=None):
        self.loader = loader
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostBatchPrefetcher stages batches onto a CUDA device (iswm_b200 has no CPU path)")
        self.image_dtype, self.label_dtype = image_dtype, label_dtype
        self._stream = torch.cuda.Stream(self.device)
        self._bufs = [None, None]                       # two device staging buffers, kept across epochs
        self.h2d_bytes = 0                              # bytes copied so far (bench.py reports them per step)

    @staticmethod
    def _split(batch):
        if isinstance(batch, dict):
            img = batch.get("image", batch.get("img"))
            return img, batch["mask"]
        if isinstance(batch, (list, tuple)) and len(batch) == 2:
            return batch[0], batch[1]
        raise ValueError(f"Unexpected batch format: {type(batch)}")

    def _stage(self, batch, slot: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Copy one host batch into the fixed device buffers of `slot` (0/1) on the side stream."""
        img, lab = self._split(batch)
        ldt = self.label_dtype or lab.dtype
        buf = self._bufs[slot]
        if buf is None or buf[0].shape != img.shape or buf[1].shape != lab.shape or buf[1].dtype != ldt:
            # (re)allocated on the consumer's stream pool, outside the steady state: no allocator churn per step
            buf = (torch.empty(img.shape, dtype=self.image_dtype, device=self.device),
                   torch.empty(lab.shape, dtype=ldt, device=self.device))
            self._bufs[slot] = buf
            self._stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._stream):
            buf[0].copy_(img, non_blocking=True)        # dtype conversion (if any) happens in the copy
            buf[1].copy_(lab, non_blocking=True)
        self.h2d_bytes += img.numel() * img.element_size() + lab.numel() * lab.element_size()
        return buf

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Yields views of two alternating device buffers: a yielded batch is valid until the batch after the
        next one is requested (double buffering), which is what a training loop needs."""
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it), 0)
        except StopIteration:
            return
        i = 0
        while nxt is not None:
            cur = nxt
            consumer = torch.cuda.current_stream(self.device)
            consumer.wait_stream(self._stream)          # the copy of THIS batch has landed
            # the other buffer was read by the previous step, already enqueued on the consumer stream:
            # the next copy may overwrite it only after that work
            self._stream.wait_event(consumer.record_event())
            try:
                nxt = self._stage(next(it), (i + 1) & 1)  # overlaps with the step the caller runs on `cur`
            except StopIteration:
                nxt = None
            i += 1
            yield cur


class DeviceTransform:
    """The tensor half of the reference's transform pipelines on the GPU, over uint8 tiles.

    val   (train.py:364-368): ExtToTensor + ExtNormalize                      -> DeviceTransform(mean, std)
    train (train.py:355-362): ... + ExtRandomCrop + ExtRandomHorizontalFlip   -> DeviceTransform(mean, std, crop_size=S, hflip=True)

    `__call__(images_u8 [B,Hs,Ws,3], labels_u8 [B,Hs,Ws]) -> (float32 [B,3,H,W], uint8 [B,H,W])` in two kernel
    launches; the host ships uint8 tiles (1/4 of the fp32 bytes, 1/8 of the int64 label bytes over PCIe) and the
    labels stay uint8 (the criterion and the metric read them as they are). Crop origins and flip decisions are drawn
    on the host from `generator` (one origin / one coin per image, as ExtRandomCrop.get_params and
    ExtRandomHorizontalFlip do per sample, utils/ext_transforms.py:350-365, :94-111). Normalisation is bit-identical
    to torchvision's to_tensor + normalize.

    The FULL train pipeline of train.py:355-362 - ExtRandomScale((0.5, 2.0)) first, ExtRandomCrop(pad_if_needed=True) -
    is `DeviceTransform(mean, std, crop_size=S, hflip=True, scale_range=(0.5, 2.0), pad_if_needed=True)`: one
    `ops.random_scale_crop` call (3 launches) that resamples ONLY the crop window of the scaled tile, bit-identical to
    Pillow's bilinear (image) / nearest (label) resize, zero padding included (utils/ext_transforms.py:94-111, :377-391).
    `draw_scaled` makes the per-sample records on the host (scale ~ U(lo, hi), target = (int(h*s), int(w*s)), padding,
    crop origin in the padded tile, flip) exactly as the reference's classes derive them from their random numbers."""

    def __init__(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), crop_size=None, hflip: bool = False,
                 p_flip: float = 0.5, generator: torch.Generator = None, scale_range=None, pad_if_needed: bool = False):
        self.mean, self.std = tuple(mean), tuple(std)
        self.crop_size = (crop_size, crop_size) if isinstance(crop_size, int) else crop_size
        self.hflip, self.p_flip = hflip, p_flip
        self.generator = generator
        self.scale_range, self.pad_if_needed = scale_range, pad_if_needed
        if scale_range is not None and self.crop_size is None:
            raise ValueError("scale_range needs crop_size: scaled tiles of one batch differ in size (the reference crops them too)")
        self._tables = None

    @staticmethod
    def scaled_geometry(Hs: int, Ws: int, scale: float, crop_hw, pad_if_needed: bool = True):
        """(sh, sw, pad, Hp, Wp) of one sample: ExtRandomScale's target size (utils/ext_transforms.py:107) and the padding
        ExtRandomCrop(pad_if_needed) adds on EVERY side (:377-385: width first, then height on the already padded tile)."""
        sh, sw = int(Hs * scale), int(Ws * scale)
        th, tw = crop_hw
        pad = 0
        if pad_if_needed and sw + 2 * pad < tw:
            pad += int((1 + tw - (sw + 2 * pad)) / 2)
        if pad_if_needed and sh + 2 * pad < th:
            pad += int((1 + th - (sh + 2 * pad)) / 2)
        return sh, sw, pad, sh + 2 * pad, sw + 2 * pad

    def draw_scaled(self, B: int, Hs: int, Ws: int, scales=None) -> torch.Tensor:
        """Host-side records of one batch for the scaled pipeline: int32 [B,8] = (sh, sw, pad, y0, x0, flip, 0, 0).
        `scales` overrides the uniform draw (tests)."""
        H, W = self.crop_size
        lo, hi = self.scale_range
        geom = torch.zeros((B, 8), dtype=torch.int32)
        for b in range(B):
            s = float(scales[b]) if scales is not None else lo + (hi - lo) * float(torch.rand((), generator=self.generator, dtype=torch.float64))
            sh, sw, pad, Hp, Wp = self.scaled_geometry(Hs, Ws, s, (H, W), self.pad_if_needed)
            if sh < 1 or sw < 1:
                raise ValueError(f"scale {s} collapses the {Hs}x{Ws} tile")
            if Hp < H or Wp < W:
                raise ValueError(f"scaled tile {sh}x{sw} smaller than the {H}x{W} crop (pad_if_needed=False)")
            if Hp == H and Wp == W:
                y0 = x0 = 0                                    # ExtRandomCrop.get_params returns the origin without drawing
            else:
                y0 = int(torch.randint(0, Hp - H + 1, (), generator=self.generator))
                x0 = int(torch.randint(0, Wp - W + 1, (), generator=self.generator))
            fl = int(float(torch.rand((), generator=self.generator)) < self.p_flip) if self.hflip else 0
            geom[b, 0], geom[b, 1], geom[b, 2], geom[b, 3], geom[b, 4], geom[b, 5] = sh, sw, pad, y0, x0, fl
        return geom

    def _call_scaled(self, images, labels, geom):
        from . import ops
        B, Hs, Ws, Cc = images.shape
        if B == 0:                                             # an empty batch (drop_last=False loaders can end on one): nothing to launch
            H, W = self.crop_size
            x = torch.empty((0, Cc, H, W), dtype=torch.float32, device=images.device)
            return x if labels is None else (x, torch.empty((0, H, W), dtype=torch.uint8, device=images.device))
        geom = geom if geom is not None else self.draw_scaled(B, Hs, Ws)
        kmax = ops.random_scale_kmax(Hs, Ws, geom.tolist())
        tab_hw = (int(geom[:, 0].max()), int(geom[:, 1].max()))
        x, y, self._tables = ops.random_scale_crop(images, labels, geom.to(images.device, non_blocking=True), self.crop_size, self.mean, self.std,
                                                   kmax, tab_hw, self._tables)
        return x if labels is None else (x, y)

    def draw(self, B: int, Hs: int, Ws: int):
        """Host-side random parameters of one batch: (origin_xy int32 [B,2] or None, flip uint8 [B] or None)."""
        org = flip = None
        if self.crop_size is not None:
            H, W = self.crop_size
            if H > Hs or W > Ws:
                raise ValueError(f"crop {H}x{W} larger than the {Hs}x{Ws} tile (pad on the host first)")
            ys = torch.randint(0, Hs - H + 1, (B,), generator=self.generator)
            xs = torch.randint(0, Ws - W + 1, (B,), generator=self.generator)
            org = torch.stack([xs, ys], 1).to(torch.int32)
        if self.hflip:
            flip = (torch.rand(B, generator=self.generator) < self.p_flip).to(torch.uint8)
        return org, flip

    def __call__(self, images: torch.Tensor, labels: torch.Tensor = None, params=None):
        from . import ops
        if not images.is_cuda:
            raise RuntimeError("DeviceTransform runs on CUDA tensors (iswm_b200 has no CPU path)")
        if self.scale_range is not None:
            return self._call_scaled(images, labels, params)
        B, Hs, Ws, _ = images.shape
        org, flip = params if params is not None else self.draw(B, Hs, Ws)
        dev = images.device
        org_d = None if org is None else org.to(dev, non_blocking=True)
        flip_d = None if flip is None else flip.to(dev, non_blocking=True)
        size = self.crop_size if self.crop_size is not None else (Hs, Ws)
        x = ops.u8_to_f32_norm(images, self.mean, self.std, size, org_d, flip_d)
        if labels is None:
            return x
        if org_d is None and flip_d is None:
            return x, labels
        return x, ops.crop_flip_u8(labels, size, org_d, flip_d)
