import tensorflow as tf

constant_initializer = tf.constant_initializer
