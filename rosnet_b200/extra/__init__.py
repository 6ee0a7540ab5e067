"""Optional integrations (reference ``rosnet/extra``): registered only if the third-party
package is importable."""
try:  # autoray is not a dependency; keep ``like="rosnet_b200"`` working when it is installed
    from autoray import autoray as _ar

    from ..dispatch import to_numpy as _to_numpy

    for _name in ("rosnet_b200", "rosnet"):
        _ar._FUNCS[_name, "to_numpy"] = _to_numpy
except Exception:  # pragma: no cover - autoray absent
    pass
