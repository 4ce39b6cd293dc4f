from .protocolBase import ProtocolBase
from .loadProtocol import loadProtocol
