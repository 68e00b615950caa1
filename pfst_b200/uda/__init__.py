from .pfgst import PFGST
from .uda_decorator import UDADecorator, get_module

__all__ = ["PFGST", "UDADecorator", "get_module"]
