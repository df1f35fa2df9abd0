from torch.utils.data import Dataset


class BaseDataset(Dataset):
    """Common ctor of the datasets: `data_dir` (Path) and `type` in {'train', 'valid', 'test'}
    (reference src/data/datasets/base_dataset.py:5-14)."""

    def __init__(self, data_dir, type):
        super().__init__()
        self.data_dir = data_dir
        self.type = type
