"""ORACLE SUPPORT: tf.contrib stand-in (layers.fully_connected, losses, slim)."""
import numpy as np

import tensorflow as tf


class _Layers(object):
    @staticmethod
    def xavier_initializer(seed=None, **kwargs):
        return ('xavier', seed)

    @staticmethod
    def variance_scaling_initializer(**kwargs):
        return ('variance_scaling', kwargs)

    @staticmethod
    def fully_connected(inputs, num_outputs, activation_fn=None, weights_initializer=None,
                        weights_regularizer=None, biases_initializer=None, scope=None,
                        **kwargs):
        """y = activation_fn(inputs @ weights[in,out] + biases); variables
        `<scope>/weights`, `<scope>/biases`, default scope `fully_connected`."""
        if activation_fn is None and 'activation_fn' not in kwargs:
            pass
        with tf.variable_scope(scope or tf.unique_default_scope('fully_connected')):
            prefix = tf.current_scope()
            w = tf.get_param(prefix + '/weights')
            b = tf.get_param(prefix + '/biases')
        x = tf._v(inputs)
        assert w.shape == (x.shape[1], num_outputs), (prefix, w.shape, x.shape, num_outputs)
        y = (x @ w + b).astype(np.float32)
        out = tf.T(y)
        return activation_fn(out) if activation_fn is not None else out

    @staticmethod
    def flatten(x):
        v = tf._v(x)
        return tf.T(v.reshape(v.shape[0], -1))


class _Losses(object):
    @staticmethod
    def add_loss(loss):
        tf._losses.append(loss)


layers = _Layers()
losses = _Losses()
