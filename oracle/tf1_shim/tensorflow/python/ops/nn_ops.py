import tensorflow as tf

bias_add = tf.nn.bias_add
