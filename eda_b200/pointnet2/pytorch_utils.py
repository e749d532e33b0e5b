"""Layer containers with the parameter / buffer names of the reference's pointnet2/pytorch_utils.py
(SharedMLP 11-36, _BNBase 39-64, _ConvBase 67-120, Conv1d/Conv2d 123-189, FC 226-261,
BNMomentumScheduler 271-296), so that reference checkpoints load unchanged:

    <mlp>.layer{i}.conv.weight                        (Cout, Cin, 1, 1), no bias when bn=True
    <mlp>.layer{i}.bn.bn.{weight,bias,running_mean,running_var,num_batches_tracked}

These classes only hold parameters and define the plain-PyTorch composition; the B200 path reads
the parameters out of them and runs the fused CUDA kernels (pointnet2_modules.py)."""
import torch.nn as nn


class BatchNorm1d(nn.Sequential):
    def __init__(self, in_size, *, name=""):
        super().__init__()
        self.add_module(name + "bn", nn.BatchNorm1d(in_size))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0.0)


class BatchNorm2d(nn.Sequential):
    def __init__(self, in_size, name=""):
        super().__init__()
        self.add_module(name + "bn", nn.BatchNorm2d(in_size))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0.0)


class _ConvBlock(nn.Sequential):
    """conv -> [bn] -> [activation]  (or bn -> activation -> conv when preact)."""

    def __init__(self, conv_cls, bn_cls, in_size, out_size, *, kernel_size, stride, padding, activation, bn, init,
                 bias, preact, name):
        super().__init__()
        use_bias = bias and not bn  # a BatchNorm right after the conv makes its bias redundant
        conv = conv_cls(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, bias=use_bias)
        init(conv.weight)
        if use_bias:
            nn.init.constant_(conv.bias, 0.0)
        norm = bn_cls(in_size if preact else out_size) if bn else None
        if preact:
            if norm is not None:
                self.add_module(name + "bn", norm)
            if activation is not None:
                self.add_module(name + "activation", activation)
        self.add_module(name + "conv", conv)
        if not preact:
            if norm is not None:
                self.add_module(name + "bn", norm)
            if activation is not None:
                self.add_module(name + "activation", activation)


class Conv1d(_ConvBlock):
    def __init__(self, in_size, out_size, *, kernel_size=1, stride=1, padding=0, activation=nn.ReLU(inplace=True),
                 bn=False, init=nn.init.kaiming_normal_, bias=True, preact=False, name=""):
        super().__init__(nn.Conv1d, BatchNorm1d, in_size, out_size, kernel_size=kernel_size, stride=stride,
                         padding=padding, activation=activation, bn=bn, init=init, bias=bias, preact=preact, name=name)


class Conv2d(_ConvBlock):
    def __init__(self, in_size, out_size, *, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_, bias=True, preact=False,
                 name=""):
        super().__init__(nn.Conv2d, BatchNorm2d, in_size, out_size, kernel_size=kernel_size, stride=stride,
                         padding=padding, activation=activation, bn=bn, init=init, bias=bias, preact=preact, name=name)


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d blocks named layer0, layer1, ... (pytorch_utils.py:11-36)."""

    def __init__(self, args, *, bn=False, activation=nn.ReLU(inplace=True), preact=False, first=False, name=""):
        super().__init__()
        for i in range(len(args) - 1):
            plain_first = first and preact and i == 0
            self.add_module(
                name + "layer{}".format(i),
                Conv2d(args[i], args[i + 1], bn=bn and not plain_first, activation=None if plain_first else activation,
                       preact=preact))

    def fusable_layers(self):
        """[(conv, bn_or_None)] if every block is conv(1x1, stride 1) -> [BatchNorm2d] -> ReLU, else None.
        The fused kernels (eda_sa_mlp_forward) implement exactly that block."""
        out = []
        for block in self.children():
            mods = dict(block.named_children())
            conv = mods.get("conv")
            act = mods.get("activation")
            norm = mods.get("bn")
            if conv is None or not isinstance(act, nn.ReLU) or list(mods)[0] != "conv":
                return None
            if tuple(conv.kernel_size) != (1, 1) or tuple(conv.stride) != (1, 1) or tuple(conv.padding) != (0, 0):
                return None
            bn = norm[0] if norm is not None else None
            if bn is not None and (not bn.affine or not bn.track_running_stats):
                return None
            out.append((conv, bn))
        return out


class FC(nn.Sequential):
    def __init__(self, in_size, out_size, *, activation=nn.ReLU(inplace=True), bn=False, init=None, preact=False,
                 name=""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0.0)
        if preact:
            if bn:
                self.add_module(name + "bn", BatchNorm1d(in_size))
            if activation is not None:
                self.add_module(name + "activation", activation)
        self.add_module(name + "fc", fc)
        if not preact:
            if bn:
                self.add_module(name + "bn", BatchNorm1d(out_size))
            if activation is not None:
                self.add_module(name + "activation", activation)


def set_bn_momentum_default(bn_momentum):
    def fn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.momentum = bn_momentum

    return fn


class BNMomentumScheduler(object):
    def __init__(self, model, bn_lambda, last_epoch=-1, setter=set_bn_momentum_default):
        if not isinstance(model, nn.Module):
            raise RuntimeError("Class '{}' is not a PyTorch nn Module".format(type(model).__name__))
        self.model = model
        self.setter = setter
        self.lmbd = bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
        self.last_epoch = epoch
        self.model.apply(self.setter(self.lmbd(epoch)))
