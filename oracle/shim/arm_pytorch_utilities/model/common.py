"""LearnableParameterizedModel restated: checkpoint naming / save / load (SURVEY.md §5 'Checkpoint')."""
import abc
import os
import glob
import torch


class LearnableParameterizedModel(abc.ABC):
    def __init__(self, root_dir, name='', lr=1e-3):
        self.root_dir = root_dir
        self.name = name
        self.step = 0
        params = []
        for m in self.modules().values():
            params += list(m.parameters())
        self.optimizer = torch.optim.Adam(params, lr=lr) if params else None

    @abc.abstractmethod
    def modules(self):
        return {}

    def parameters(self):
        ps = []
        for m in self.modules().values():
            ps += list(m.parameters())
        return ps

    def device(self):
        # used at hybrid_model.py:197 (use_temp_local_nominal_model) and controller.py:162
        return next(iter(self.parameters())).device

    def eval(self):
        for m in self.modules().values():
            m.eval()

    def train(self):
        for m in self.modules().values():
            m.train()

    def get_last_checkpoint(self, sort_by_time=True):
        base = os.path.join(self.root_dir, 'checkpoints')
        cands = [f for f in glob.glob(os.path.join(base, glob.escape(self.name) + '.*.tar'))
                 if f.rsplit('.', 2)[-2].isdigit()]
        if not cands:
            return None
        if sort_by_time:
            return max(cands, key=os.path.getctime)
        return max(cands, key=lambda f: int(f.rsplit('.', 2)[-2]))

    def save(self, last=False):
        state = {'step': self.step, 'state_dict': {k: m.state_dict() for k, m in self.modules().items()},
                 'optimizer': self.optimizer.state_dict()}
        base = os.path.join(self.root_dir, 'checkpoints')
        os.makedirs(base, exist_ok=True)
        torch.save(state, os.path.join(base, '{}.{}.tar'.format(self.name, self.step)))

    def load(self, filename):
        if not filename or not os.path.isfile(filename):
            return False
        ckpt = torch.load(filename, map_location='cpu', weights_only=False)
        self.step = ckpt['step']
        for k, m in self.modules().items():
            m.load_state_dict(ckpt['state_dict'][k])
        return True
