"""CPU: the C-ABI library builds, loads, and exports every symbol that
include/gossipnet_b200.h declares; the ctypes table covers all of them.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

from gossipnet_b200 import _lib, build
from tests.helpers import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'gossipnet_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gn_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ('gn_iou_dense', 'gn_neighbor_count', 'gn_neighbor_fill', 'gn_pwfeat_mlp_fwd',
                 'gn_block_pair_fwd', 'gn_fc_fwd', 'gn_detection_matching', 'gn_loss_fwd'):
        assert must in syms


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.gn_abi_version() >= 1


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.gn_last_error() is not None
