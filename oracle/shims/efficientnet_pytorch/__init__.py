"""Stand-in for efficientnet_pytorch==0.7.1 so that the reference's terrain_encoder/lss.py imports.

TEST INFRASTRUCTURE ONLY.  The real package is not under /root/reference and cannot be installed
(no network); its `from_pretrained` would also download ImageNet weights.  The architecture is
restated once, in monoforce_b200/efficientnet.py (state_dict-compatible, parameter count pinned
to B0's 5,288,548); this shim re-exports it with `from_pretrained` = random initialisation.
Consequence: everything lss.py does AROUND the trunk (Up, depthnet, depth softmax x features,
frustum geometry, voxel pooling, BevEncode, heads) is pinned against the unmodified reference;
the trunk's internals are "parity unpinned".
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from monoforce_b200.efficientnet import EfficientNet  # noqa: E402,F401
