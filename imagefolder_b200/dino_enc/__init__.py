from .dinov2 import DINOv2Decoder, DINOv2Encoder
from .to_pixel import ToPixel
from .vision_transformer import VisionTransformer, create_model

__all__ = ["DINOv2Encoder", "DINOv2Decoder", "ToPixel", "VisionTransformer", "create_model"]
