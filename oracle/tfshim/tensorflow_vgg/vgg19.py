"""ORACLE / test infrastructure: stand-in for the base class the reference's `custom_vgg19.py` derives from,
`tensorflow_vgg.vgg19.Vgg19` of machrisaa/tensorflow-vgg (README.md:39; un-vendored, unpinned: "master").  Only the
three members `custom_Vgg19` uses are restated, from that repository's published vgg19.py:

    conv_layer(bottom, name) = relu(bias_add(conv2d(bottom, filter, strides 1, padding SAME), biases))   NHWC
    avg_pool(bottom, name)   = tf.nn.avg_pool(ksize 2x2, strides 2x2, padding SAME)                        NHWC
    get_conv_filter / get_bias: tf.constant(data_dict[name][0 | 1])        filter [3,3,Cin,Cout], bias [Cout]

With this module on the path the reference's OWN custom_vgg19.py (input scaling, BGR mean subtraction, layer order)
and the Gram branches of its loss.py run unmodified on oracle/tfshim."""
import tensorflow as tf


class Vgg19:
    def __init__(self, vgg19_npy_path=None):
        self.data_dict = None

    def avg_pool(self, bottom, name):
        return tf.nn.avg_pool(bottom, ksize=[1, 2, 2, 1], strides=[1, 2, 2, 1], padding='SAME', name=name)

    def conv_layer(self, bottom, name):
        with tf.variable_scope(name):
            filt = self.get_conv_filter(name)
            conv = tf.nn.conv2d(bottom, filt, [1, 1, 1, 1], padding='SAME')
            conv_biases = self.get_bias(name)
            bias = tf.nn.bias_add(conv, conv_biases)
            return tf.nn.relu(bias)

    def get_conv_filter(self, name):
        return tf.constant(self.data_dict[name][0], name='filter')

    def get_bias(self, name):
        return tf.constant(self.data_dict[name][1], name='biases')
