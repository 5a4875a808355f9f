"""``HyperparametersMixin`` as the reference uses it (``.hparams`` attribute dict,
``save_hyperparameters()`` capturing the constructor arguments; reference
utils/lightning.py:18-80 serialises it).  Lightning's own mixin is used when the
package is installed so that checkpoints round-trip; otherwise a local equivalent."""
from __future__ import annotations

import inspect

try:  # pragma: no cover - depends on the environment
    from pytorch_lightning.core.mixins.hparams_mixin import HyperparametersMixin  # type: ignore
except Exception:  # noqa: BLE001

    class AttributeDict(dict):
        def __getattr__(self, key):
            try:
                return self[key]
            except KeyError as e:  # deepcopy / pickle probe dunder attributes
                raise AttributeError(key) from e

        def __setattr__(self, key, value):
            self[key] = value

    class HyperparametersMixin:
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)

        @property
        def hparams(self) -> AttributeDict:
            if "_hparams" not in self.__dict__:
                self.__dict__["_hparams"] = AttributeDict()
            return self.__dict__["_hparams"]

        def save_hyperparameters(self, *args, ignore=None, frame=None, logger=True) -> None:
            ignore = {ignore} if isinstance(ignore, str) else set(ignore or ())
            if args and isinstance(args[0], dict):
                init_args = dict(args[0])
            else:
                frame = frame or inspect.currentframe().f_back
                info = inspect.getargvalues(frame)
                init_args = {}
                for name in info.args:
                    if name != "self":
                        init_args[name] = info.locals[name]
                if info.keywords:
                    init_args.update(info.locals[info.keywords])
                init_args.pop("__class__", None)
                if args:
                    init_args = {k: v for k, v in init_args.items() if k in args}
            for k, v in init_args.items():
                if k not in ignore:
                    self.hparams[k] = v
