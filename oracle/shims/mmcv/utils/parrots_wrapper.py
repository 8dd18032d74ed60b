from torch.nn.modules.batchnorm import _BatchNorm  # noqa
from torch.nn.modules.instancenorm import _InstanceNorm  # noqa
