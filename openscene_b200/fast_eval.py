"""The fused engine behind the UNMODIFIED call site ``model(sinput)`` (run/evaluate.py:289, :307).

The reference drives the network module by module (models/mink_unet.py:116-174): ~200 Python-level operator calls per
scene, each a kernel launch plus torch allocations -- 2.25x slower than `engine.FusedMinkUNet` on the same kernels.  This
module gives the reference's own call the engine's speed without an edit to ``run/evaluate.py``:

1. ``ME.SparseTensor(feats, coords)`` created under ``torch.no_grad()`` arms a (temporary) global forward pre-hook that
   watches which ``nn.Module`` receives that tensor.  The innermost module that gets it before the first operator of this
   package runs -- ``MinkUNet34C`` inside ``DisNet`` inside an optional DDP wrapper -- is the network root.
2. The root's ``forward`` is wrapped (instance attribute; class and state dict untouched).  Eligible calls -- eval mode,
   grad disabled, a fresh input tensor, no extra arguments -- go to a ``FusedMinkUNet`` built from the root; everything
   else (training, autograd, anything unusual) goes to the original forward.
3. Nothing is taken on trust: the first eligible call runs BOTH paths and installs the engine only if the per-point
   relative error is below 1e-3; a root that is not a BasicBlock MinkUNet, or whose outputs differ, keeps the
   module-by-module path for good (``root._osb_fast.disabled`` says why).  Parameter / buffer updates
   (``load_state_dict``, fine-tuning steps, ``.to()``) are detected through ``_version`` / ``data_ptr`` and trigger a
   re-fold + re-validation.

``OSB_FAST_EVAL=0`` switches the mechanism off.
"""
import os
import weakref

import torch

_ENABLED = os.environ.get('OSB_FAST_EVAL', '1') != '0'
_state = {'handle': None, 'watch': None, 'candidate': None}


def enabled():
    return _ENABLED


def set_enabled(flag):
    global _ENABLED
    _ENABLED = bool(flag)
    if not _ENABLED:
        _disarm()


def watch(x):
    """called by SparseTensor.__init__ (grad disabled): find the module that consumes `x`"""
    if not _ENABLED:
        return
    _state['watch'] = weakref.ref(x)
    _state['candidate'] = None
    if _state['handle'] is None:
        _state['handle'] = torch.nn.modules.module.register_module_forward_pre_hook(_pre_hook)


def _disarm():
    h = _state['handle']
    if h is not None:
        h.remove()
    _state['handle'] = _state['watch'] = _state['candidate'] = None


def _pre_hook(module, args):
    w = _state['watch']
    x = w() if w is not None else None
    if x is None:
        _disarm()
        return None
    if not args or args[0] is not x:
        return None
    if getattr(module, '_osb_me_op', False):                   # first operator reached: the last candidate is the root
        root = _state['candidate']
        _disarm()
        if root is not None and not hasattr(root, '_osb_fast'):
            install(root)
        return None
    if hasattr(module, '_osb_fast'):                           # steady state: the root is already wrapped
        _disarm()
        return None
    _state['candidate'] = module
    return None


def install(root):
    ff = FastForward(root, root.forward)
    root.__dict__['_osb_fast'] = ff
    root.__dict__['forward'] = ff                              # instance attribute: nn.Module.__call__ picks it up
    return ff


def uninstall(root):
    root.__dict__.pop('forward', None)
    root.__dict__.pop('_osb_fast', None)


class FastForward:
    """callable standing in for ``root.forward``"""

    def __init__(self, root, orig):
        self.root = weakref.ref(root)
        self.orig = orig
        self.engine = None
        self.key = None
        self.validated = False
        self.disabled = None                                   # reason string once the fast path is given up
        self.calls_fast = 0
        self.last_err = None
        self._tensors = None

    def __deepcopy__(self, memo):                              # copies of the model start without a wrapper
        return None

    def _key(self, root):
        if self._tensors is None:                              # cached: walking the module tree costs more than the check
            self._tensors = list(root.parameters()) + list(root.buffers())
        return tuple((t._version, t.data_ptr()) for t in self._tensors)

    def __call__(self, x, *args, **kwargs):
        root = self.root()
        from .me import SparseTensor
        if (self.disabled is not None or not _ENABLED or args or kwargs or root is None or root.training
                or torch.is_grad_enabled() or type(x) is not SparseTensor or not x._is_fresh_input()):
            return self.orig(x, *args, **kwargs)
        key = self._key(root)
        if self.engine is None or key != self.key:
            from .engine import FusedMinkUNet
            try:
                self.engine = FusedMinkUNet(root)
            except (NotImplementedError, AttributeError, RuntimeError, TypeError) as e:
                self.disabled = f'not a fusable MinkUNet: {e}'
                return self.orig(x)
            self.key, self.validated = key, False
        if not self.validated:
            ref = self.orig(x)
            if not (torch.is_tensor(ref) and ref.dim() == 2 and ref.shape[0] == x._F_ext.shape[0]):
                self.disabled = 'forward does not return the [N, C] feature matrix'
                return ref
            out = self.engine(x._raw_coords, x._F_ext)
            if out.shape != ref.shape:
                self.disabled = f'shape mismatch {tuple(out.shape)} vs {tuple(ref.shape)}'
                return ref
            err = float(((out - ref).norm(dim=1) / (ref.norm(dim=1) + 1e-20)).max())
            self.last_err = err
            if err < 1e-3:
                self.validated = True
            else:
                self.disabled = f'engine and module path differ (max per-point relative error {err:.3e})'
            return ref
        self.calls_fast += 1
        out = self.engine(x._raw_coords, x._F_ext)
        x._cm = self.engine.last_cm                            # later uses of the input tensor share the manager
        return out
