"""ToPixel head -- reference tokenizer/tokenizer_image/dino_enc/to_pixel.py:36-95 ('linear' and
'identity'; the 'conv'/'siren' variants are not selected by any shipped config)."""
import torch
import torch.nn as nn


class ToPixel(nn.Module):
    def __init__(self, to_pixel='linear', img_size=256, in_channels=3, in_dim=512, patch_size=16) -> None:
        super().__init__()
        self.to_pixel_name = to_pixel
        self.patch_size = patch_size
        self.num_patches = (img_size // patch_size) ** 2
        self.in_channels = in_channels
        if to_pixel == 'linear':
            self.model = nn.Linear(in_dim, in_channels * patch_size * patch_size)
        elif to_pixel == 'identity':
            self.model = nn.Identity()
        else:
            raise NotImplementedError(f"to_pixel={to_pixel!r}: only 'linear' / 'identity' are built")

    def get_last_layer(self):
        return self.model.weight if self.to_pixel_name == 'linear' else None

    def unpatchify(self, x):
        """x: (N, L, patch_size**2 * 3) -> imgs: (N, 3, H, W)        (to_pixel.py:70-81)"""
        p = self.patch_size
        h = w = int(x.shape[1] ** .5)
        assert h * w == x.shape[1]
        x = x.reshape(shape=(x.shape[0], h, w, p, p, 3))
        x = torch.einsum('nhwpqc->nchpwq', x)
        return x.reshape(shape=(x.shape[0], 3, h * p, h * p))

    def forward(self, x):
        if self.to_pixel_name == 'linear':
            x = self.unpatchify(self.model(x))
        return x
