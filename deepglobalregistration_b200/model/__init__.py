"""Model registry with the reference's ``load_model(name)`` contract
(model/__init__.py:24-38): class lookup by name, ``None`` for unknown names."""
import logging

from . import resunet

MODELS = [getattr(resunet, a) for a in dir(resunet) if 'Net' in a]


def load_model(name):
  table = {m.__name__: m for m in MODELS}
  if name not in table:
    logging.info(f'Invalid model index. You put {name}. Options are: {sorted(table)}')
    return None
  return table[name]
