"""Build libgossipnet_b200.so in-tree with nvcc for sm_100a (no torch headers)."""
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc')
LIB_PATH = os.path.join(CSRC, 'libgossipnet_b200.so')


def build(verbose=False, jobs=8):
    cmd = ['make', '-C', CSRC, '-j%d' % jobs]
    if not verbose:
        cmd.append('-s')
    subprocess.check_call(cmd)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('build did not produce ' + LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    print(build(verbose=True))
