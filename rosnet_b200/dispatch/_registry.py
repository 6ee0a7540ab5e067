"""A small type-dispatch registry (the product's own; ``multimethod`` is not a dependency).

The reference declares one ``multimethod`` per NumPy function name whose default raises
``NotImplementedError`` (``rosnet/dispatch/numpy.py:4-68``) and lets array types register
overloads keyed on their annotations.  ``Dispatcher`` gives the same surface:

    tensordot = Dispatcher("tensordot")
    @tensordot.register
    def _(a: BlockArray, b: BlockArray, axes): ...
    tensordot(x, y, axes)        # picks the most specific matching overload
    tensordot[(BlockArray, BlockArray)]   # fetch an overload by types

An annotation may be a class (``isinstance``; structural ABCs work), ``Sequence[T]`` (a
list/tuple whose first element matches ``T``), a ``Union`` or missing (matches anything).
"""
import collections.abc
import inspect
import typing


def _score(arg, ann):
    if ann is inspect.Parameter.empty or ann is typing.Any:
        return 0
    origin = typing.get_origin(ann)
    if origin is typing.Union:
        best = None
        for alt in typing.get_args(ann):
            s = _score(arg, alt)
            if s is not None and (best is None or s > best):
                best = s
        return best
    if origin in (collections.abc.Sequence, list, tuple):
        if not isinstance(arg, (list, tuple)):
            return None
        inner = typing.get_args(ann)
        if not inner or not arg:
            return 1
        s = _score(arg[0], inner[0])
        return None if s is None else s + 1
    if isinstance(ann, type):
        return 2 if isinstance(arg, ann) else None
    return 0


class Dispatcher:
    def __init__(self, name: str):
        self.__name__ = name
        self._overloads = []  # (annotations, fn, number of positional parameters, takes *args)

    def _annotations(self, fn):
        try:
            hints = typing.get_type_hints(fn)
        except Exception:
            hints = getattr(fn, "__annotations__", {})
        params = inspect.signature(fn).parameters.values()
        return tuple(
            hints.get(p.name, inspect.Parameter.empty)
            for p in params
            if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)
        )

    def _add(self, anns, fn):
        # the signature is inspected once, here: __call__ sits on the np.tensordot route of every contraction step
        params = list(inspect.signature(fn).parameters.values())
        npos = sum(p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) for p in params)
        self._overloads.append((anns, fn, npos, any(p.kind == p.VAR_POSITIONAL for p in params)))

    def register(self, *args):
        """``@d.register`` (types from annotations) or ``@d.register(T1, T2, ...)``."""
        if len(args) == 1 and inspect.isfunction(args[0]):
            self._add(self._annotations(args[0]), args[0])
            return args[0]

        def deco(fn):
            self._add(tuple(args) if args else self._annotations(fn), fn)
            return fn

        return deco

    def _resolve(self, args):
        best, best_score = None, -1
        for anns, fn, npos, varargs in self._overloads:
            if len(args) > npos and not varargs:
                continue
            total = 0
            for a, ann in zip(args, anns):
                s = _score(a, ann)
                if s is None:
                    total = None
                    break
                total += s
            if total is not None and total > best_score:
                best, best_score = fn, total
        return best

    def __call__(self, *args, **kwargs):
        fn = self._resolve(args)
        if fn is None:
            kinds = ", ".join(type(a).__name__ for a in args)
            raise NotImplementedError(f"{self.__name__} has no overload for ({kinds})")
        return fn(*args, **kwargs)

    def __getitem__(self, types_):
        if not isinstance(types_, tuple):
            types_ = (types_,)
        for anns, fn, _npos, _varargs in reversed(self._overloads):
            ok = True
            for t, ann in zip(types_, anns):
                base = typing.get_origin(ann) or ann
                if ann is inspect.Parameter.empty or not isinstance(base, type):
                    continue
                if not (isinstance(t, type) and issubclass(t, base)):
                    ok = False
                    break
            if ok:
                return fn
        raise KeyError(types_)
