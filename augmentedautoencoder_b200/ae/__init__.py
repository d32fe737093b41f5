from . import ae_factory as factory  # same alias as auto_pose/ae/__init__.py:1
from . import utils
from .session import Session, placeholder, variable_scope
