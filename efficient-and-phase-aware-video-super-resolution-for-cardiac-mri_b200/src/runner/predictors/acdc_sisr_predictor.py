"""Test loop of the single-image nets (reference src/runner/predictors/acdc_sisr_predictor.py:14-170).

Per frame: SR = net(lr), the configured losses, PSNR / SSIM (/ Cardiac*) on the de-normalised images; optional export
of a per-frame CSV, one PNG per frame and one GIF per slice.  B200 differences (results identical): frames of equal
shape are batched into one launch (`frames_per_launch`; the reference runs one frame per iteration), frames are
sharded over the ranks under torch.distributed, and the scalars of a launch cross PCIe once.
"""
import csv
import functools
import logging
from pathlib import Path

import torch

from pvsr import parallel
from src.utils import denormalize
from .acdc_vsr_refinenet_predictor import AcdcVSRRefineNetPredictor, _write_png
from .base_predictor import BasePredictor, per_sample_scores


class AcdcSISRPredictor(BasePredictor):
    dataset_name = 'acdc'

    def __init__(self, saved_dir=None, exported=False, frames_per_launch=32, **kwargs):
        super().__init__(**kwargs)
        if self.test_dataloader.batch_size != 1:
            raise ValueError(f'The testing batch size should be 1. Got {self.test_dataloader.batch_size}.')
        self.exported = exported
        if exported:
            self.saved_dir = Path(saved_dir)
        self.frames_per_launch = max(1, int(frames_per_launch))
        self._denormalize = functools.partial(denormalize, dataset=self.dataset_name)

    def _outputs(self, lr):
        return [self.net(lr)]

    def _name(self, index):
        filename = Path(self.test_dataloader.dataset.data[index][0]).parts[-1].split('.')[0]
        patient, _, sid, fid = filename.split('_')
        return filename, patient, sid, fid

    def _my_indices(self, n_items):
        """Frames of this rank: whole (patient, slice) groups, cost-balanced by frame count, so that every slice's
        GIF (and its PNGs) is written by exactly one rank from ALL of that slice's frames."""
        if self.world <= 1:
            return list(range(n_items))
        groups = {}
        for index in range(n_items):
            _, patient, sid, _ = self._name(index)
            groups.setdefault((patient, sid), []).append(index)
        keys = sorted(groups)
        owned = parallel.shard_indices(len(keys), self.rank, self.world, sizes=[len(groups[k]) for k in keys])
        return sorted(i for g in owned for i in groups[keys[g]])

    def predict(self):
        self.net.eval()
        dataset = self.test_dataloader.dataset
        mine = self._my_indices(len(dataset))
        names = [fn.__class__.__name__ for fn in self.metric_fns + self.loss_fns]
        log, count, rows, frames = self._init_log(), 0, [], {}

        def flush(items):
            nonlocal count
            lr = torch.stack([it['lr_img'] for _, it in items]).to(self.device, non_blocking=True)
            hr = torch.stack([it['hr_img'] for _, it in items]).to(self.device, non_blocking=True)
            with torch.no_grad():
                outs = self._outputs(lr)             # [net(lr)]; the SRFB variant: the num_steps outputs
                sr = outs[-1]
                srd, hrd = self._denormalize(sr), self._denormalize(hr)
                patients = [self._name(index)[1] for index, _ in items]
                losses, metrics = per_sample_scores(self.loss_fns, self.metric_fns, sr, hr, srd, hrd, patients,
                                                    dataset=self.dataset_name)
                if len(outs) > 1:                    # losses averaged over the steps (acdc_sisr_srfb_predictor.py:103-107)
                    per_step = [losses] + [per_sample_scores(self.loss_fns, [], o, hr, None, None, patients,
                                                             dataset=self.dataset_name)[0] for o in outs[:-1]]
                    losses = torch.stack(per_step).mean(dim=0)
                flat = torch.cat([metrics, losses], dim=1).cpu()
                imgs = srd[:, 0].to(torch.uint8).cpu().numpy() if self.exported else None
            nm = len(self.metric_fns)
            for n, (index, _) in enumerate(items):
                metrics, losses = flat[n, :nm], flat[n, nm:]
                log['Loss'] += float((losses * self.loss_weights.cpu()).sum())
                for fn, v in zip(self.loss_fns, losses.tolist()):
                    log[fn.__class__.__name__] += v
                for fn, v in zip(self.metric_fns, metrics.tolist()):
                    log[fn.__class__.__name__] += v
                count += 1
                if self.exported:
                    filename, patient, sid, fid = self._name(index)
                    rows.append([filename, *metrics.tolist(), *losses.tolist()])
                    idir = self.saved_dir / 'imgs' / patient
                    idir.mkdir(parents=True, exist_ok=True)
                    _write_png(idir / f'{sid}_{fid}.png', imgs[n])
                    frames.setdefault((patient, sid), []).append((fid, imgs[n]))

        pending = {}
        for index in mine:
            item = dataset[index]
            key = tuple(item['lr_img'].shape)
            pending.setdefault(key, []).append((index, item))
            if len(pending[key]) == self.frames_per_launch:
                flush(pending.pop(key))
        for items in pending.values():
            flush(items)

        log, count = parallel.reduce_log(log, count, self.device)
        if self.exported:
            for (patient, sid), fs in frames.items():          # one GIF per slice (reference :70-77)
                vdir = self.saved_dir / 'videos' / patient
                vdir.mkdir(parents=True, exist_ok=True)
                AcdcVSRRefineNetPredictor._dump_video(self, vdir / (sid.replace('slice', 'sequence') + '.gif'),
                                                      [img for _, img in sorted(fs, key=lambda f: f[0])])
            rows = AcdcVSRRefineNetPredictor._gather_rows(self, rows)
            if self.rank == 0:
                self.saved_dir.mkdir(parents=True, exist_ok=True)
                with open(self.saved_dir / 'results.csv', 'w', newline='') as f:
                    csv.writer(f).writerows([['name'] + names] + sorted(rows, key=lambda r: r[0]))
        log = {k: v / max(count, 1) for k, v in log.items()}
        logging.info(f'Test log: {log}.')
        return log


class Dsb15SISRPredictor(AcdcSISRPredictor):
    dataset_name = 'dsb15'


class AcdcSISRSRFBPredictor(AcdcSISRPredictor):
    """Iterated single-image nets - DRFSISRNet on this path (reference acdc_sisr_srfb_predictor.py:13-127): losses are
    averaged over the `num_steps` outputs, metrics and exports use the last one."""

    def _outputs(self, lr):
        return list(self.net(lr))


class Dsb15SISRSRFBPredictor(AcdcSISRSRFBPredictor):
    dataset_name = 'dsb15'
