import torch


class Data(object):
    """PyG-1.x style attribute bag: `.keys` is a property, `__cat_dim__(key, value)` takes 2 args."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k in self.__dict__.keys() if not k.startswith('__') and self.__dict__[k] is not None]

    def __getitem__(self, key):
        return getattr(self, key, None)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.keys

    def __cat_dim__(self, key, value):
        return -1 if ('index' in key or 'face' in key) and key != 'bbox_idx' else 0

    @property
    def num_nodes(self):
        for k in ('x', 'pos'):
            v = getattr(self, k, None)
            if v is not None:
                return v.size(0)
        return None


class InMemoryDataset(torch.utils.data.Dataset):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root = root
        self.transform = transform


def extract_zip(*a, **k):
    raise NotImplementedError
