"""Parameter containers whose state_dict keys follow a (key -> shape) manifest.

The product keeps the reference's checkpoint format without restating its module classes: a manifest
(manifest.py) is expanded into nested, otherwise empty nn.Module nodes with parameters / buffers at the leaves.
"""
import torch
from torch import nn

from .manifest import is_int_key


class _Node(nn.Module):
    pass


class ParamTree(nn.Module):
    def build_tree(self, manifest, init=None):
        """Register every manifest entry on `self` (nested nodes for dotted keys)."""
        self._manifest = dict(manifest)
        for key, shape in manifest.items():
            node = self
            parts = key.split(".")
            for part in parts[:-1]:
                if part not in node._modules:
                    node.add_module(part, _Node())
                node = node._modules[part]
            leaf = parts[-1]
            if is_int_key(key):
                node.register_buffer(leaf, torch.zeros(tuple(shape), dtype=torch.long))
            else:
                val = init[key] if init is not None and key in init else torch.zeros(tuple(shape))
                node.register_parameter(leaf, nn.Parameter(val.clone().float(), requires_grad=False))
        return self

    def flat_state_dict(self):
        """name -> tensor (no copy, no prefix) restricted to the manifest keys."""
        sd = nn.Module.state_dict(self)
        return {k: sd[k] for k in self._manifest}
