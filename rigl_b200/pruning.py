"""Registry of masked layers -- the stand-in for the graph collections of
`tensorflow.contrib.model_pruning.python.pruning` that the reference reads
through `pruning.get_masks() / get_weights() / get_masked_weights()`
(rigl/sparse_optimizers.py:46-56; collection names visible at
rigl/mnist/mnist_train_eval.py:236).  Entries are returned in creation order.
"""


class MaskedLayerRegistry(object):

  def __init__(self):
    self._layers = []

  def register(self, layer):
    if layer not in self._layers:
      self._layers.append(layer)

  def clear(self):
    del self._layers[:]

  def layers(self):
    return list(self._layers)

  def get_masks(self):
    return [l.mask for l in self._layers]

  def get_weights(self):
    return [l.weight for l in self._layers]

  def get_masked_weights(self):
    return [l.masked_weights for l in self._layers]

  @classmethod
  def from_module(cls, module):
    """Registry holding the masked layers found in `module`, in module order."""
    reg = cls()
    for m in module.modules():
      if getattr(m, 'is_rigl_masked_layer', False):
        reg.register(m)
    return reg


_DEFAULT = MaskedLayerRegistry()


def default_registry():
  return _DEFAULT


def reset_default_registry():
  """Analogue of tf.reset_default_graph() for the mask collections."""
  _DEFAULT.clear()


def get_masks():
  return _DEFAULT.get_masks()


def get_weights():
  return _DEFAULT.get_weights()


def get_masked_weights():
  return _DEFAULT.get_masked_weights()
