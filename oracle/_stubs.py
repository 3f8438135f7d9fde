"""Import-only stand-ins for modules the reference imports but the MPPI path never executes
(pybullet, tensorboardX, matplotlib, seaborn, gym, ROS, and the arm_pytorch_utilities
submodules that are off the path).  TEST INFRASTRUCTURE ONLY.

A stub module answers any attribute with a subclassable, callable placeholder class, so statements like
`class MixingBatchGP(gpytorch.models.ExactGP)` or `from arm_pytorch_utilities.gmm import GMM` import;
calling into a placeholder at run time raises, so nothing on the measured/parity path can silently
depend on one."""
import importlib.abc
import importlib.machinery
import sys
import types

STUB_ROOTS = (
    'tensorboardX', 'pybullet', 'pybullet_data', 'matplotlib', 'seaborn', 'gym', 'rospy',
    'std_msgs', 'geometry_msgs', 'visualization_msgs', 'window_recorder', 'arm_video_recorder', 'mujoco_py',
    'tampc_or', 'tampc_or_msgs',
)
STUB_SUBMODULES = tuple('arm_pytorch_utilities.' + n for n in (
    'simulation', 'grad', 'gmm', 'trajectory', 'policy', 'draw', 'softknn', 'serialization'))


class _PlaceholderMeta(type):
    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _make_placeholder('{}.{}'.format(cls.__name__, name))

    def __iter__(cls):
        return iter(())


def _make_placeholder(qualname):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError('oracle shim: placeholder {} was called; it is not on the MPPI path'.format(qualname))

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        raise RuntimeError('oracle shim: placeholder {}.{} was used at run time'.format(qualname, name))

    return _PlaceholderMeta(qualname, (), {'__init__': __init__, '__call__': __call__, '__getattr__': __getattr__})


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        full = '{}.{}'.format(self.__name__, name)
        if full in sys.modules:
            return sys.modules[full]
        obj = _make_placeholder(full)
        setattr(self, name, obj)
        return obj


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split('.')[0]
        if root in STUB_ROOTS or fullname in STUB_SUBMODULES or \
                any(fullname.startswith(s + '.') for s in STUB_SUBMODULES):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    global _installed
    if not _installed:
        sys.meta_path.append(_StubFinder())
        _installed = True
