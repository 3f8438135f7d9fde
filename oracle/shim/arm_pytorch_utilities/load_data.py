"""load_data restated: DataConfig (checkpoint-name token), DataLoader over .mat files (env/env.py:12)."""
import os
import numpy as np
import scipy.io
import torch
import torch.utils.data


class DataConfig:
    def __init__(self, sort_data=True, predict_difference=True, predict_all_dims=True, force_affine=False,
                 expanded_input=False, y_in_x_space=True):
        self.sort_data = sort_data
        self.predict_difference = predict_difference
        self.predict_all_dims = predict_all_dims
        self.force_affine = force_affine
        self.expanded_input = expanded_input
        self.y_in_x_space = y_in_x_space
        self.nx = None
        self.nu = None
        self.ny = None
        self.n_input = None

    def load_data_info(self, x, u=None, y=None, full_input=None):
        self.nx = x.shape[1]
        self.nu = u.shape[1] if u is not None else 0
        self.ny = y.shape[1] if y is not None else self.nx
        self.n_input = full_input.shape[1] if full_input is not None else self.nx + self.nu

    def input_dim(self):
        if not self.n_input:
            self.n_input = self.nx + (self.nu or 0)
        return self.n_input

    def options(self):
        return (self.sort_data, self.predict_difference, self.predict_all_dims, self.force_affine,
                self.expanded_input, self.y_in_x_space)

    def __str__(self):
        return "i{}_o{}_s{}_pd{}_pa{}_a{}_e{}_y{}".format(self.input_dim(), self.ny,
                                                          *(int(c) for c in self.options()))

    __repr__ = __str__


class DataLoader:
    """Loads one .mat file (or every .mat in a directory) under file_cfg.DATA_DIR through the subclass's
    `_process_file_raw_data(d) -> (xu, y, info)`."""

    def __init__(self, file_cfg=None, dir_load_callback=None, config=None):
        # the reference passes its `tampc.cfg` MODULE (env/env.py:13-16); keep only the path so that a datasource
        # holding this loader stays deep-copyable (hybrid_model.py:131 deep-copies it for every local model)
        self.data_dir = getattr(file_cfg, 'DATA_DIR', None)
        self.config = config if config is not None else DataConfig()
        self.dir_load_callback = dir_load_callback

    def _process_file_raw_data(self, d):
        raise NotImplementedError

    def load_file(self, full_filename):
        raw = scipy.io.loadmat(full_filename)
        return self._process_file_raw_data(raw)

    def load(self, dir_or_file):
        full = os.path.join(self.data_dir, dir_or_file)
        if os.path.isfile(full):
            return self.load_file(full)
        parts = []
        for root, _, files in sorted(os.walk(full)):
            for f in sorted(files):
                if f.endswith('.mat'):
                    parts.append(self.load_file(os.path.join(root, f)))
        return tuple(np.row_stack([p[i] for p in parts]) for i in range(len(parts[0])))


def split_train_validation(data, validation_ratio):
    """Deterministic split (upstream shuffles with the global numpy seed; unpinned)."""
    n = data[0].shape[0]
    n_val = int(round(n * validation_ratio))
    perm = np.random.RandomState(0).permutation(n)
    val_idx = np.sort(perm[:n_val])
    train_idx = np.sort(perm[n_val:])
    return tuple(d[train_idx] for d in data), tuple(d[val_idx] for d in data)


class SimpleDataset(torch.utils.data.Dataset):
    def __init__(self, *seqs):
        self.seqs = [s for s in seqs if s is not None]

    def __len__(self):
        return len(self.seqs[0])

    def __getitem__(self, i):
        return tuple(s[i] for s in self.seqs)
