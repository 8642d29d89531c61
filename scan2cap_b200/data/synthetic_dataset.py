"""A torch Dataset with the item contract of the reference's ScannetReferenceDataset (lib/dataset.py:338-362 channel
order, :503-538 key set / dtypes) backed by the seeded procedural scenes of scan2cap_b200/synthetic.py -- ScanNet /
ScanRefer are licensed and not shipped with the reference, so this is what lets scripts/train.py-style loops run end
to end.  Exposes the attributes the reference's solver and model construction read: ``vocabulary``, ``glove``
(embeddings), ``weights`` (lib/solver.py:300, scripts/train.py:63-64).  Items are numpy arrays per key; the default
collate stacks them, pin_memory gives the pinned batches TrainStep.prefetch() copies from."""
import numpy as np
import torch
from torch.utils.data import Dataset

from .. import synthetic
from .scannet.model_util_scannet import ScannetDatasetConfig


class SyntheticScan2CapDataset(Dataset):
    def __init__(self, num_scenes=64, num_points=40000, use_height=True, use_color=False, use_normal=False,
                 use_multiview=False, num_vocabs=3500, seed=42, mean_size_arr=None, lang_len=None):
        assert not use_color, "the procedural scenes carry no colour"
        self.num_scenes, self.num_points = num_scenes, num_points
        self.use_height, self.use_normal, self.use_multiview = use_height, use_normal, use_multiview
        self.num_vocabs, self.seed, self.lang_len = num_vocabs, seed, lang_len
        self.mean_size_arr = ScannetDatasetConfig().mean_size_arr if mean_size_arr is None else mean_size_arr
        self.vocabulary, self.glove, _ = synthetic.make_vocabulary(num_vocabs, seed=seed)
        self.weights = np.ones(num_vocabs, np.float32)   # (the reference's class weights; unused by get_scene_cap_loss)

    def __len__(self):
        return self.num_scenes

    def __getitem__(self, idx):
        d = synthetic.make_data_dict(1, self.num_points, use_normal=self.use_normal, use_multiview=self.use_multiview,
                                     use_height=self.use_height, num_vocabs=self.num_vocabs,
                                     seed=self.seed + 7919 * int(idx), mean_size_arr=self.mean_size_arr,
                                     lang_len=self.lang_len)
        item = {k: v[0] for k, v in d.items()}
        item["dataset_idx"] = np.array(idx).astype(np.int64)
        item["load_time"] = np.float64(0.0)
        return item


def to_pinned(batch):
    """Pin every tensor of a collated batch (DataLoader(pin_memory=True) does the same in its worker thread)."""
    return {k: (v.pin_memory() if isinstance(v, torch.Tensor) and not v.is_pinned() else v) for k, v in batch.items()}
