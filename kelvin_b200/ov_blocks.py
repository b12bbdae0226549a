"""Attribute containers for occupied/virtual integral blocks.

Same names and constructor signatures as the (un-vendored) cqcpy.ov_blocks
classes the reference builds at kelvin/cc_utils.py:588,599-601,769-775, so that
``F.ov``, ``I.vvoo``, ``Iabab.ovvo`` ... read the same in both code bases.
Block values are CUDA float64 tensors (NumPy arrays are accepted and moved on
first use by the kernels' wrappers).
"""


class one_e_blocks(object):
    def __init__(self, oo, ov, vo, vv):
        self.oo = oo
        self.ov = ov
        self.vo = vo
        self.vv = vv


class two_e_blocks(object):
    names = ("vvvv", "vvvo", "vovv", "vvoo", "vovo", "oovv", "vooo", "ooov", "oooo")

    def __init__(self, vvvv=None, vvvo=None, vovv=None, vvoo=None, vovo=None,
                 oovv=None, vooo=None, ooov=None, oooo=None):
        self.vvvv = vvvv
        self.vvvo = vvvo
        self.vovv = vovv
        self.vvoo = vvoo
        self.vovo = vovo
        self.oovv = oovv
        self.vooo = vooo
        self.ooov = ooov
        self.oooo = oooo


class two_e_blocks_full(object):
    names = ("vvvv", "vvvo", "vvov", "vovv", "ovvv", "vvoo", "vovo", "ovvo",
             "voov", "ovov", "oovv", "vooo", "ovoo", "oovo", "ooov", "oooo")

    def __init__(self, vvvv=None, vvvo=None, vvov=None, vovv=None, ovvv=None, vvoo=None,
                 vovo=None, ovvo=None, voov=None, ovov=None, oovv=None, vooo=None,
                 ovoo=None, oovo=None, ooov=None, oooo=None):
        loc = locals()
        for nm in self.names:
            setattr(self, nm, loc[nm])
