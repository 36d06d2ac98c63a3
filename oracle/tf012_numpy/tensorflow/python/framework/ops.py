"""ORACLE SUPPORT: decorators the reference's roi_pooling_op_grad.py applies."""


def RegisterGradient(name):
    return lambda fn: fn
