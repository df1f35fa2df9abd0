"""Test loop of RefineNet (reference src/runner/predictors/acdc_vsr_refinenet_predictor.py:15-183).

Per cine sequence: SR frames = the LAST output list of the net (:62), per-frame L1 and PSNR/SSIM(/Cardiac*) on the
de-normalised frames, optional export of a per-frame CSV, one GIF per sequence and one PNG per frame.

B200 differences (results identical):
  * only the consumed output list is computed (`net.only_last_head`), the other 3S-1 heads are dead work;
  * sequences of equal shape are batched into one plan launch (`sequences_per_launch`; the reference processes one
    sequence per iteration and requires batch_size 1);
  * under torch.distributed the sequences are sharded over the ranks by cost (frames x pixels); metric sums are
    all-reduced, CSV rows gathered on rank 0;
  * per-frame scalars come back in one device->host transfer per launch instead of one `.item()` each.
"""
import csv
import functools
import logging
from pathlib import Path

import numpy as np
import torch

from pvsr import parallel
from src.utils import denormalize
from .base_predictor import BasePredictor, per_sample_scores


def _write_png(path, img):
    """8-bit grayscale PNG without imageio / scipy.misc (neither is installed here)."""
    import struct
    import zlib
    h, w = img.shape
    raw = b''.join(b'\x00' + img[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack('>I', len(data)) + tag + data
        return c + struct.pack('>I', zlib.crc32(tag + data) & 0xffffffff)

    with open(path, 'wb') as f:
        f.write(b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, 0, 0, 0, 0)) +
                chunk(b'IDAT', zlib.compress(raw, 6)) + chunk(b'IEND', b''))


class AcdcVSRRefineNetPredictor(BasePredictor):
    dataset_name = 'acdc'

    def __init__(self, saved_dir=None, exported=False, sequences_per_launch=8, **kwargs):
        super().__init__(**kwargs)
        if self.test_dataloader.batch_size != 1:
            raise ValueError(f'The testing batch size should be 1. Got {self.test_dataloader.batch_size}.')
        self.exported = exported
        if exported:
            self.saved_dir = Path(saved_dir)
        self.sequences_per_launch = max(1, int(sequences_per_launch))
        self._denormalize = functools.partial(denormalize, dataset=self.dataset_name)

    # ------------------------------------------------------------------ main loop
    def predict(self):
        self.net.eval()
        if hasattr(self.net, 'only_last_head'):
            self.net.only_last_head = True
        dataset = self.test_dataloader.dataset
        collate = self.test_dataloader.collate_fn
        mine = parallel.shard_indices(len(dataset), self.rank, self.world)
        header = ['name'] + [fn.__class__.__name__ for fn in self.metric_fns + self.loss_fns]
        rows = []
        log, count = self._init_log(), 0

        # a DeviceDataloader cuts batches out of HBM-resident volumes itself (pvsr/device_loader.py)
        device_side = hasattr(self.test_dataloader, 'fetch')
        pending = {}    # shape -> list of (index, item)
        def flush(items):
            nonlocal count
            idx = [i for i, _ in items]
            batch = (self.test_dataloader.fetch(idx) if device_side
                     else self._allocate_data(collate([it for _, it in items])))
            inputs, targets, pos_codes, _ = self._get_inputs_targets(batch)
            with torch.no_grad():
                outputs = self._forward(inputs, pos_codes)          # T x (n, 1, H, W)
                per_seq = self._per_sequence(outputs, targets, idx)
            for n, (index, losses, metrics, sr) in enumerate(per_seq):
                T = losses.shape[0]
                loss = (losses.mean(dim=0) * self.loss_weights.cpu()).sum()
                self._update_log(log, 1, T, loss, losses, metrics)
                count += T
                if self.exported:
                    rows.extend(self._export(index, losses, metrics, sr))

        for index in mine:
            if device_side:
                item, key = None, self.test_dataloader.item_key(index)
            else:
                item = dataset[index]
                key = (len(item['lr_imgs']),) + tuple(item['lr_imgs'][0].shape)
            pending.setdefault(key, []).append((index, item))
            if len(pending[key]) == self.sequences_per_launch:
                flush(pending.pop(key))
        for items in pending.values():
            flush(items)

        log, count = parallel.reduce_log(log, count, self.device)
        if self.exported:
            rows = self._gather_rows(rows)
            if self.rank == 0:
                self.saved_dir.mkdir(parents=True, exist_ok=True)
                with open(self.saved_dir / 'results.csv', 'w', newline='') as f:
                    csv.writer(f).writerows([header] + sorted(rows, key=lambda r: r[0]))
        log = {k: v / max(count, 1) for k, v in log.items()}
        logging.info(f'Test log: {log}.')
        return log

    def _forward(self, inputs, pos_codes):
        """The SR frames the scores are computed on: the last output list (:62)."""
        return self.net(inputs, pos_codes)[-1]

    def _get_inputs_targets(self, batch):
        return batch['lr_imgs'], batch['hr_imgs'], batch['pos_code'], batch['index']

    # ------------------------------------------------------------------ per-sequence numbers
    def _patient(self, index):
        name = Path(self.test_dataloader.dataset.data[index][0]).parts[-1].split('.')[0]
        patient, _, sid = name.split('_')
        return name, patient, sid

    def _per_sequence(self, outputs, targets, indices):
        """[(index, losses (T, #loss), metrics (T, #metric), sr uint8 (T, H, W) or None)] for every sequence of the
        launch; every loss / metric scores all n * T frames in one call and all scalars cross PCIe once."""
        n, T = outputs[0].shape[0], len(outputs)
        out = torch.stack(outputs, dim=1).flatten(0, 1)          # (n * T, 1, H, W), sequence-major
        tgt = torch.stack(targets, dim=1).flatten(0, 1)
        patients = [self._patient(indices[s])[1] for s in range(n) for _ in range(T)]
        losses, metrics = per_sample_scores(self.loss_fns, self.metric_fns, out, tgt, lambda: self._denormalize(out),
                                            lambda: self._denormalize(tgt), patients, dataset=self.dataset_name)
        nl = losses.shape[1]
        flat = torch.cat([losses, metrics], dim=1).cpu().view(n, T, -1)
        frames = None
        if self.exported:
            srd = self._denormalize(out)
            frames = srd.view(n, T, *srd.shape[1:])[:, :, 0].to(torch.uint8).cpu().numpy()
        return [(indices[s], flat[s, :, :nl], flat[s, :, nl:], None if frames is None else frames[s])
                for s in range(n)]

    def _compute_losses(self, outputs, targets):
        """(T, #loss_fns) per-frame losses of one sequence (reference :123-136)."""
        return torch.stack([torch.stack([fn(o, t) for o, t in zip(outputs, targets)]) for fn in self.loss_fns], dim=1)

    def _compute_metrics(self, outputs, targets, name):
        """(T, #metric_fns) per-frame metrics of one sequence on de-normalised frames (reference :138-158)."""
        sr = [self._denormalize(o) for o in outputs]
        hr = [self._denormalize(t) for t in targets]
        cols = []
        for fn in self.metric_fns:
            extra = (name,) if 'Cardiac' in fn.__class__.__name__ else ()
            cols.append(torch.stack([fn(o, t, *extra) for o, t in zip(sr, hr)]))
        return torch.stack(cols, dim=1)

    def _update_log(self, log, batch_size, T, loss, losses, metrics):
        w = batch_size * T
        log['Loss'] += float(loss) * w
        for fn, v in zip(self.loss_fns, losses.mean(dim=0).tolist()):
            log[fn.__class__.__name__] += v * w
        for fn, v in zip(self.metric_fns, metrics.mean(dim=0).tolist()):
            log[fn.__class__.__name__] += v * w

    # ------------------------------------------------------------------ export
    def _export(self, index, losses, metrics, sr):
        name, patient, sid = self._patient(index)
        name = name.replace('2d+1d', '2d').replace('sequence', 'slice')
        rows = [[f'{name}_frame{t + 1:0>2d}', *metrics[t].tolist(), *losses[t].tolist()] for t in range(len(sr))]
        vdir, idir = self.saved_dir / 'videos' / patient, self.saved_dir / 'imgs' / patient
        vdir.mkdir(parents=True, exist_ok=True)
        idir.mkdir(parents=True, exist_ok=True)
        self._dump_video(vdir / f'{sid}.gif', sr)
        stem = sid.replace('sequence', 'slice')
        for t, img in enumerate(sr):
            _write_png(idir / f'{stem}_frame{t + 1:0>2d}.png', img)
        return rows

    def _dump_video(self, path, imgs):
        """Animated GIF of the SR frames; falls back to an .npy stack when no GIF writer is installed."""
        try:
            import imageio
            with imageio.get_writer(path) as writer:
                for img in imgs:
                    writer.append_data(img)
        except ImportError:
            try:
                from PIL import Image
                frames = [Image.fromarray(img) for img in imgs]
                frames[0].save(path, save_all=True, append_images=frames[1:], loop=0)
            except ImportError:
                np.save(Path(path).with_suffix('.npy'), np.stack(imgs))

    def _gather_rows(self, rows):
        if not parallel.is_distributed():
            return rows
        import torch.distributed as dist
        gathered = [None] * self.world if self.rank == 0 else None
        dist.gather_object(rows, gathered, dst=0)
        return [r for part in gathered for r in part] if self.rank == 0 else []


class Dsb15VSRRefineNetPredictor(AcdcVSRRefineNetPredictor):
    """Same loop with the DSB15 de-normalisation constants (src/utils.py:15-16)."""
    dataset_name = 'dsb15'
