"""rectools_b200: a B200-native (sm_100a) scoring + top-K engine for RecTools' vector-ranking path.

Public surface (host-side mirror of `rectools.models.rank`):
  * `Distance`, `B200Ranker`              -- drop-in for `ImplicitRanker` / the `Ranker` protocol
  * `B200TorchRanker`                     -- `TorchRanker`-signature adapter (transformer id-embedding scorers)
  * `install()` / `uninstall()`           -- rebind the ranker used by `VectorModel` / `EASEModel` in an installed rectools
  * `recommend()`                         -- vectorised `ModelBase.recommend` around the ranker (cached viewed CSR, id maps, table)
  * `ShardedB200Ranker`                   -- item-sharded multi-GPU ranking (one process per GPU, NCCL all-gather + merge)
The CUDA library is `rectools_b200/libb200rank.so` (C ABI: include/b200_rank.h); build it with
`python -m rectools_b200.build`.  There is no CPU fallback.
"""
from .ranker import B200Ranker, Distance, Engine, flatten_padded  # noqa: F401
from .integration import B200ImplicitRanker, B200TorchRanker, install, uninstall  # noqa: F401
from .recommend import recommend, recommend_to_items  # noqa: F401

__all__ = [
    "B200Ranker",
    "B200ImplicitRanker",
    "B200TorchRanker",
    "Distance",
    "Engine",
    "flatten_padded",
    "install",
    "recommend",
    "recommend_to_items",
    "uninstall",
]
__version__ = "0.1.0"
