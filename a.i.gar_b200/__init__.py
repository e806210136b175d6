"""B200-native batched agar.io env step (host side).  See DESIGN.md.

`aigar_b200` is the importable alias of this directory (its literal name contains dots)."""
from .layout import (AgarConfig, AgarLayout, Record, derive_config, layout_for_config,  # noqa: F401
                     BOT_NN, BOT_GREEDY, BOT_RANDOM, OBS_REFERENCE, OBS_CANONICAL)


def __getattr__(name):  # lazy: importing the package must not require a GPU or torch
    if name in ("AgarBatch", "BatchedModel", "load_library"):
        from . import env as _env
        return getattr(_env, name)
    raise AttributeError(name)
