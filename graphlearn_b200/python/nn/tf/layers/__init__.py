"""graphlearn/python/nn/tf/layers/*: one module per layer in the reference; the names resolve here as well
(``from graphlearn.python.nn.tf.layers.sage_conv import SAGEConv``)."""
