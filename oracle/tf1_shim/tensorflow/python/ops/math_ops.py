import tensorflow as tf

matmul = tf.matmul
