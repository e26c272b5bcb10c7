import tensorflow as tf

variable_scope = tf.variable_scope
get_variable_scope = tf.get_variable_scope
get_variable = tf.get_variable
