"""`flax.core`: FrozenDict / freeze / unfreeze."""
import collections.abc


class FrozenDict(collections.abc.Mapping):
    def __init__(self, *args, **kwargs):
        self._d = {k: (FrozenDict(v) if isinstance(v, dict) else v) for k, v in dict(*args, **kwargs).items()}

    def __getitem__(self, k):
        return self._d[k]

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __hash__(self):
        return hash(tuple(sorted(self._d)))

    def __repr__(self):
        return f"FrozenDict({self._d!r})"


def freeze(d):
    return FrozenDict(d)


def unfreeze(d):
    if isinstance(d, (FrozenDict, dict)):
        return {k: unfreeze(v) for k, v in d.items()}
    return d
