from .BaselineModel import BaselineModel  # noqa: F401
from .DyYOLO import DyYOLO  # noqa: F401
