"""ORACLE SUPPORT: tf.contrib.slim is imported by the reference's network.py
but only used by the ResNet image-feature variant (out of scope)."""
