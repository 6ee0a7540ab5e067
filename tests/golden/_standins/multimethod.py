"""Throw-away stand-in for the PyPI `multimethod` package (absent here, no network).

GENERATOR SCAFFOLDING ONLY: it exists so that tests/golden/make_golden.py can import the
unmodified reference from /root/reference and record golden vectors.  It is not part of
the product and nothing under rosnet_b200/ imports it.
"""
import collections.abc
import inspect
import types
import typing


def _match(arg, ann):
    """Score how well `arg` fits annotation `ann`; None = no match, larger = more specific."""
    if ann is inspect.Parameter.empty or ann is typing.Any:
        return 0
    origin = typing.get_origin(ann)
    if origin is typing.Union:
        scores = [s for s in (_match(arg, a) for a in typing.get_args(ann)) if s is not None]
        return max(scores) if scores else None
    if origin is typing.Literal:
        return 3 if any(arg == v or arg is v for v in typing.get_args(ann)) else None
    if origin in (collections.abc.Sequence, list, tuple, typing.Sequence):
        if isinstance(arg, (str, bytes)) or not isinstance(arg, (list, tuple)):
            return None
        inner = typing.get_args(ann)
        if not inner or len(arg) == 0:
            return 1
        s = _match(arg[0], inner[0])
        return None if s is None else s + 1
    if origin is not None and isinstance(origin, type):
        if not isinstance(arg, origin):
            return None
        params = typing.get_args(ann)
        oc = getattr(arg, "__orig_class__", None)
        if params and oc is not None:
            have = typing.get_args(oc)
            for p, h in zip(params, have):
                if isinstance(p, type) and isinstance(h, type) and not issubclass(h, p):
                    return None
            return 3
        return 2
    if isinstance(ann, type):
        try:
            return 2 if isinstance(arg, ann) else None
        except TypeError:
            return None
    return 0


class multimethod:
    def __init__(self, fn):
        self.__name__ = getattr(fn, "__name__", "multimethod")
        self.__doc__ = getattr(fn, "__doc__", None)
        self.default = fn
        self.impls = []  # (annotations tuple, fn)
        sig = inspect.signature(fn)
        anns = [p.annotation for p in sig.parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        if any(a is not inspect.Parameter.empty for a in anns):
            self._add(fn, None)

    def _add(self, fn, explicit):
        if explicit is None:
            hints = {}
            try:
                hints = typing.get_type_hints(fn)
            except Exception:
                hints = getattr(fn, "__annotations__", {})
            sig = inspect.signature(fn)
            anns = tuple(
                hints.get(p.name, inspect.Parameter.empty)
                for p in sig.parameters.values()
                if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)
            )
        else:
            anns = tuple(explicit)
        self.impls.append((anns, fn))

    def register(self, *args):
        if len(args) == 1 and inspect.isfunction(args[0]):
            fn = args[0]
            self._add(fn, None)
            return self if getattr(fn, "__name__", None) == self.__name__ else fn

        def deco(fn):
            self._add(fn, args if args else None)
            return self if getattr(fn, "__name__", None) == self.__name__ else fn

        return deco

    def _pick(self, args):
        best, best_score = None, None
        for anns, fn in self.impls:
            sig = inspect.signature(fn)
            nreq = sum(
                1
                for p in sig.parameters.values()
                if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.default is p.empty
            )
            has_var = any(p.kind == p.VAR_POSITIONAL for p in sig.parameters.values())
            npos = sum(1 for p in sig.parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD))
            if len(args) > npos and not has_var:
                continue
            score, ok = 0, True
            for a, ann in zip(args, anns):
                s = _match(a, ann)
                if s is None:
                    ok = False
                    break
                score += s
            if not ok:
                continue
            if best_score is None or score > best_score:
                best, best_score = fn, score
        return best

    def __call__(self, *args, **kwargs):
        fn = self._pick(args)
        if fn is None:
            fn = self.default
        return fn(*args, **kwargs)

    def __getitem__(self, types_):
        if not isinstance(types_, tuple):
            types_ = (types_,)
        for anns, fn in reversed(self.impls):
            ok = True
            for t, ann in zip(types_, anns):
                base = typing.get_origin(ann) or ann
                if ann is inspect.Parameter.empty:
                    continue
                if isinstance(base, type) and isinstance(t, type):
                    try:
                        if not issubclass(t, base):
                            ok = False
                    except TypeError:
                        ok = False
            if ok:
                return fn
        raise KeyError(types_)

    def __get__(self, obj, cls=None):
        if obj is None:
            return self
        return types.MethodType(self, obj)
