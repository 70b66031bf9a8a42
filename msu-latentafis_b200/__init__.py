"""latentafis_b200 — B200-native 1-vs-N gallery matcher behind the MSU-LatentAFIS matcher interface.

The directory name carries a hyphen, so the package is loaded by path (see `__graft_entry__.load_package`)
under the module name `msu_latentafis_b200`.
"""
from . import templates  # noqa: F401
from .matcher import (LafisError, Matcher, MatcherGroup, PackedGallery, PackedLatents, load_library, pack_latents,  # noqa: F401
                      pack_rolled)

__all__ = ["Matcher", "MatcherGroup", "LafisError", "PackedGallery", "PackedLatents", "pack_latents", "pack_rolled", "load_library",
           "templates"]
