"""Decoder + reconstruction loss.  Mirrors auto_pose/ae/decoder.py:13-144 (constructor arguments, ``x``,
``reconstr_loss``, ``reconstruction_target``)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .encoder import _DeviceModule
from .session import Tensor, scoped, to_device_input
from .utils import lazy_property


class Decoder(_DeviceModule):

    def __init__(self, reconstruction_target, latent_code, num_filters, kernel_size, strides, loss, bootstrap_ratio,
                 auxiliary_mask, batch_norm, is_training=False, max_batch=64, seed=43, n_encoder_convs=None, precision=None):
        if batch_norm:
            raise NotImplementedError("BATCH_NORMALIZATION: True is not supported")
        if auxiliary_mask:
            raise NotImplementedError("AUXILIARY_MASK: True is not supported (False in the template cfg)")
        if loss != "L2":
            raise NotImplementedError("LOSS: %s is not supported (template cfg uses L2)" % loss)
        L = _lib.lib()
        self._create, self._destroy = L.aae_decoder_create, L.aae_decoder_destroy
        self._set, self._get = L.aae_decoder_set_weights, L.aae_decoder_get_weights
        self._range_status = L.aae_decoder_range_status
        self._reconstruction_target = reconstruction_target
        self._latent_code = latent_code
        self._auxiliary_mask = auxiliary_mask
        self._num_filters = list(num_filters)      # already reversed by build_decoder (ae_factory.py:62)
        self._kernel_size = int(kernel_size)
        self._strides = list(strides)
        self._loss = loss
        self._bootstrap_ratio = int(bootstrap_ratio)
        self._batch_normalization = batch_norm
        self._is_training = is_training
        self.max_batch = int(max_batch)
        h, w, c = reconstruction_target.get_shape().as_list()[1:]
        self._out_shape = (h, w, c)
        latent = latent_code.get_shape().as_list()[-1]
        self._latent = latent
        nl = len(self._num_filters)
        dims = [int(h / np.prod(self._strides[i:])) for i in range(nl)]
        k0 = n_encoder_convs if n_encoder_convs is not None else nl
        var_shapes = [(scoped("dense_1/kernel"), (latent, dims[0] * dims[0] * self._num_filters[0]),
                       scoped("dense_1/bias"), (dims[0] * dims[0] * self._num_filters[0],))]
        cin = self._num_filters[0]
        for j, f in enumerate(self._num_filters[1:] + [c]):
            base = scoped("conv2d_%d" % (k0 + j))
            var_shapes.append((base + "/kernel", (self._kernel_size, self._kernel_size, cin, f), base + "/bias", (f,)))
            cin = f
        # same default as the encoder: tensor cores unless precision=_lib.PREC_FP32_SIMT is asked for
        self._auto_precision = precision is None
        if precision is None:
            precision = _lib.PREC_TC_SPLIT
        self.precision = int(precision)
        # the C ABI takes the encoder-order filters/strides and reverses them itself (aae_net_cfg)
        self._init_module((h, w, c, list(reversed(self._num_filters)), list(reversed(self._strides)), self._kernel_size,
                           latent, self.max_batch, self.precision), var_shapes, seed)
        self.reconstr_loss

    @property
    def reconstruction_target(self):
        return self._reconstruction_target

    def decode_device(self, z_dev):
        dev = z_dev.device
        h = self.handle(dev)
        B = z_dev.shape[0]
        out = torch.empty((B,) + self._out_shape, dtype=torch.float32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for a in range(0, B, self.max_batch):
            e = min(B, a + self.max_batch)
            _lib.check(_lib.lib().aae_decoder_forward(h, _lib.ptr(z_dev[a:e].contiguous()), e - a, _lib.ptr(out[a:e]), stream), "decoder forward")
        return out

    @lazy_property
    def x(self):
        def fn(ctx):
            if self not in ctx.touched:
                ctx.touched.append(self)
            return self.decode_device(to_device_input(ctx.get(self._latent_code), ctx.session.device))   # a fed latent may be numpy
        return Tensor("conv2d_out/Sigmoid", (None,) + self._out_shape, np.float32, fn)

    @staticmethod
    def loss_device(x_dev, target_dev, bootstrap_ratio, with_grad=False):
        """Bootstrapped L2 on device tensors -> (loss 0-d tensor, grad or None)."""
        if target_dev.dtype == torch.uint8:          # the kernel reads float*: a uint8 target is the image / 255 (as the trainer's feed)
            target_dev = target_dev.to(torch.float32) / 255.0
        if x_dev.dtype != torch.float32 or target_dev.dtype != torch.float32 or x_dev.shape != target_dev.shape:
            raise ValueError("bootstrapped L2 wants float32 tensors of one shape, got %s %s / %s %s"
                             % (x_dev.dtype, tuple(x_dev.shape), target_dev.dtype, tuple(target_dev.shape)))
        B = x_dev.shape[0]
        numel = x_dev[0].numel()
        loss = torch.empty((1,), dtype=torch.float32, device=x_dev.device)
        grad = torch.empty_like(x_dev) if with_grad else None
        stream = C.c_void_p(torch.cuda.current_stream(x_dev.device).cuda_stream)
        _lib.check(_lib.lib().aae_bootstrap_l2_loss(_lib.ptr(x_dev.contiguous()), _lib.ptr(target_dev.contiguous()), B, numel,
                                                    int(bootstrap_ratio), _lib.ptr(loss), _lib.ptr(grad), stream), "bootstrap_l2")
        return loss[0], grad

    @lazy_property
    def reconstr_loss(self):
        def fn(ctx):
            x = ctx.get(self.x)
            y = to_device_input(ctx.get(self._reconstruction_target), ctx.session.device)
            return self.loss_device(x, y, self._bootstrap_ratio)[0]
        return Tensor("reconstr_loss", (), np.float32, fn)
