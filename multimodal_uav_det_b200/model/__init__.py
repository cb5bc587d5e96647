from .BaselineModel import BaselineModel  # noqa: F401
from .DySOEM_SimFPN import DySOEM_SimFPN  # noqa: F401
from .DyYOLO import DyYOLO  # noqa: F401
from .RTMUAVDet import RTMUAVDet  # noqa: F401  (not exported by the reference's model/__init__.py: deprecated there)
