"""Re-export of torcheasyrec_b200/peer_exchange.py (kept so that tests/test_peer_exchange_model.py and try_peer.py
find it under the name they always used)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from torcheasyrec_b200.peer_exchange import *  # noqa: E402,F401,F403
from torcheasyrec_b200.peer_exchange import PeerState, _PeerPooled, enable_peer_exchange  # noqa: E402,F401


def declare(L):
    """ctypes signatures for a stand-alone (host-compiled) build of tzk_peer.cu."""
    from torcheasyrec_b200._lib import SIGNATURES

    for name in ("tzk_peer_pooled_gather_fwd", "tzk_peer_barrier", "tzk_peer_pull_counts", "tzk_peer_pull"):
        if hasattr(L, name):
            getattr(L, name).restype, getattr(L, name).argtypes = SIGNATURES[name]
    return L
