"""ACDCSR / DSB15SR-shaped synthetic cine sequences with the item contract of AcdcVSRRefineNetDataset (the real
datasets are not available offline).  Used by the benchmarks, the tests and `src.main` smoke configs."""
import torch

from .base_dataset import BaseDataset
from .acdc_vsr_refinenet_dataset import window_slices


class SyntheticCineDataset(BaseDataset):
    def __init__(self, downscale_factor=4, num_sequences=8, num_phases=30, lr_size=(54, 63), num_frames=7,
                 num_updated_frames=6, end_systole=11, seed=1234, data_dir=None, type='test', **_):
        super().__init__(data_dir=data_dir, type=type)
        self.downscale_factor, self.num_phases = downscale_factor, num_phases
        self.lr_size = tuple(lr_size)
        self.num_frames, self.num_updated_frames = num_frames, num_updated_frames
        self.end_systole, self.seed = end_systole, seed
        if type == 'train':
            self.data = [(f'synthetic{n:03d}_2d+1d_sequence{n:02d}', None, t) for n in range(num_sequences)
                         for t in range(num_phases)]
        else:
            self.data = [(f'synthetic{n:03d}_2d+1d_sequence{n:02d}', None) for n in range(num_sequences)]

    def __len__(self):
        return len(self.data)

    def _frames(self, seq):
        g = torch.Generator().manual_seed(self.seed + seq)
        T, (h, w), s = self.num_phases, self.lr_size, self.downscale_factor
        lr = [torch.randn(1, h, w, generator=g) for _ in range(T)]
        hr = [torch.randn(1, h * s, w * s, generator=g) for _ in range(T)]
        return lr, hr

    # ---- device-resident serving (pvsr.device_loader.DeviceDataloader) -------------------------------------------
    def sequence_table(self):
        from pvsr.synthetic import positional_code
        code = positional_code(self.num_phases, self.end_systole)
        table = []
        for seq in sorted({int(e[0][9:12]) for e in self.data}):
            lr, hr = self._frames(seq)
            table.append((torch.stack(lr, dim=-1).permute(1, 2, 0, 3).numpy(),      # (H, W, 1, T)
                          torch.stack(hr, dim=-1).permute(1, 2, 0, 3).numpy(), code))
        return table

    def transform_plan(self):
        return 0.0, 1.0, []

    def window(self, index):
        entry = self.data[index]
        train = self.type == 'train'
        return (int(entry[0][9:12]),) + window_slices(self.num_phases, entry[2] if train else 0, self.num_frames,
                                                      self.num_updated_frames, train)

    def __getitem__(self, index):
        from pvsr.synthetic import positional_code
        entry = self.data[index]
        seq = int(entry[0][9:12])
        T = self.num_phases
        lr, hr = (frames * 3 for frames in self._frames(seq))
        code = torch.from_numpy(positional_code(T, self.end_systole)).repeat(3).unsqueeze(1)
        a, b, c, d = window_slices(T, entry[2] if self.type == 'train' else 0, self.num_frames,
                                   self.num_updated_frames, self.type == 'train')
        return {'lr_imgs': lr[a:b], 'hr_imgs': hr[c:d], 'pos_code': code[a:b], 'index': index}
