from .fusion import make_fusion
from .head import make_head
from .model import PtGenerator, PtTransformerEarlyFusionIterative
from .text_net import make_text_net
from .video_net import make_video_net
