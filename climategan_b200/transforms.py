"""``climategan.transforms.DiffTransforms`` (transforms.py:609-626) — the differentiable augmentation in front of the painter
discriminator (``gen.p.diff_aug``) — as one fused kernel pass (``ops.diff_aug``) instead of ~25 ATen launches.

The per-sample random draws are made with torch ON THE TENSOR'S DEVICE, in the reference's order, shapes and dtypes (brightness,
contrast, saturation ``torch.rand(N,1,1,1)``; translation then cutout ``torch.randint(..., [N,1,1])`` rows before columns), so a
seeded run consumes the generator exactly like the reference; they are packed into the kernel's [N, 8] table without a host
round trip."""
import torch

from . import ops


class DiffTransforms:
    def __init__(self, diff_aug_opts):
        self.do_color_jittering = diff_aug_opts.do_color_jittering
        self.do_cutout = diff_aug_opts.do_cutout
        self.do_translation = diff_aug_opts.do_translation
        self.cutout_ratio = diff_aug_opts.cutout_ratio
        self.translation_ratio = diff_aug_opts.translation_ratio

    def draw(self, tensor):
        """(params [N, 8] fp32 on tensor.device, cut_h, cut_w) for one call, consuming the generator like the reference."""
        assert len(tensor.shape) == 4
        n, _, h, w = tensor.shape
        dev = tensor.device
        params = torch.zeros(n, 8, dtype=torch.float32, device=dev)
        params[:, 1:3] = 1.0
        if self.do_color_jittering:                                                     # transforms.py:493-533
            params[:, 0] = torch.rand(n, 1, 1, 1, dtype=tensor.dtype, device=dev).view(n).float() - 0.5
            params[:, 1] = torch.rand(n, 1, 1, 1, dtype=tensor.dtype, device=dev).view(n).float() + 0.5
            params[:, 2] = torch.rand(n, 1, 1, 1, dtype=tensor.dtype, device=dev).view(n).float() * 2
        if self.do_translation:                                                         # :580-606
            shift_x, shift_y = int(h * self.translation_ratio + 0.5), int(w * self.translation_ratio + 0.5)
            params[:, 3] = torch.randint(-shift_x, shift_x + 1, size=[n, 1, 1], device=dev).view(n).float()
            params[:, 4] = torch.randint(-shift_y, shift_y + 1, size=[n, 1, 1], device=dev).view(n).float()
        cut_h = cut_w = 0
        if self.do_cutout:                                                              # :546-577
            cut_h, cut_w = int(h * self.cutout_ratio + 0.5), int(w * self.cutout_ratio + 0.5)
            params[:, 5] = torch.randint(0, h + (1 - cut_h % 2), size=[n, 1, 1], device=dev).view(n).float()
            params[:, 6] = torch.randint(0, w + (1 - cut_w % 2), size=[n, 1, 1], device=dev).view(n).float()
            if cut_h == 0 or cut_w == 0:   # a ratio that rounds to an empty box cuts nothing in the reference either
                cut_h = cut_w = 0
        return params, cut_h, cut_w

    def __call__(self, tensor):
        params, cut_h, cut_w = self.draw(tensor)
        return ops.diff_aug(tensor, params, cut_h, cut_w)
