"""datasource restated: DataSource / FileDataSource (env/env.py:191, pendulum.py:31-60 show the contract)."""
import copy
import torch
from arm_pytorch_utilities import load_data, preprocess


class DataSource:
    def __init__(self, N=None, config=None, preprocessor=None, device='cpu', **kwargs):
        self.N = N
        self.config = config if config is not None else load_data.DataConfig()
        self._original_config = None
        self.preprocessor = None
        if preprocessor is not None:
            self.preprocessor = preprocessor if isinstance(preprocessor, preprocess.DatasetPreprocessor) \
                else preprocess.DatasetPreprocessor(preprocessor)
        self.d = device
        self._train = None
        self._val = None
        self._original_train = None
        self._original_val = None
        self.make_data()

    def make_data(self):
        raise NotImplementedError

    def update_preprocessor(self, preprocessor):
        """Change the preprocessor and re-make (re-fit) the data; returns the untransformed config."""
        if preprocessor is not None and not isinstance(preprocessor, preprocess.DatasetPreprocessor):
            preprocessor = preprocess.DatasetPreprocessor(preprocessor)
        self.preprocessor = preprocessor
        if self._original_config is not None:
            self.config = copy.deepcopy(self._original_config)
        self.make_data()
        return self.original_config()

    def training_set(self, original=False):
        if original and self._original_train is not None:
            return self._original_train
        return self._train

    def validation_set(self, original=False):
        if original and self._original_val is not None:
            return self._original_val
        return self._val

    def original_validation_set(self):
        return self.validation_set(original=True)

    def current_config(self):
        return self.config

    def original_config(self):
        return self._original_config if self._original_config is not None else self.config

    def data_id(self):
        return "{}".format(self.N)


class FileDataSource(DataSource):
    def __init__(self, loader, data_dir, validation_ratio=0.2, config=None, **kwargs):
        self.loader = loader
        if config is not None:
            self.loader.config = config
        self._data_dir = data_dir
        self._validation_ratio = validation_ratio
        super().__init__(config=self.loader.config, **kwargs)

    def make_data(self):
        full_set = self.loader.load(self._data_dir)
        if self._original_config is None:
            self._original_config = copy.deepcopy(self.loader.config)
            self.config = copy.deepcopy(self._original_config)
        train, val = load_data.split_train_validation(full_set, self._validation_ratio)
        to_t = lambda a: torch.from_numpy(a).to(device=self.d, dtype=torch.double)
        self._train = tuple(to_t(a) for a in train)
        self._val = tuple(to_t(a) for a in val)
        self.N = self._train[0].shape[0]
        if self.preprocessor:
            self.preprocessor.tsf.fit(*self._train)
            self.preprocessor.update_data_config(self.config)
            self._original_train = self._train
            self._original_val = self._val
            with torch.no_grad():
                self._train = self.preprocessor.tsf.transform(*self._train)
                self._val = self.preprocessor.tsf.transform(*self._val)

    def data_id(self):
        return "{}_{}".format(self._data_dir, self.N)
