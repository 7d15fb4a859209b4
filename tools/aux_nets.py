"""Stock-PyTorch stand-ins for the two auxiliary networks of the reference's generator / discriminator step
(vae_scripts/train_vae.py:268-275,298-299,395-401,449-458), for TIMING the full BASELINE configs[2] step only:

  PatchDiscriminatorRef : monai.networks.nets.PatchDiscriminator(spatial_dims=2, num_layers_d=3, channels=32, in_channels=1,
                          out_channels=1, norm="INSTANCE") restated -- 692,769 parameters, LSGAN criterion
  LPIPSSqueezeRef       : PerceptualLoss(network_type="squeeze") = LPIPS on SqueezeNet-1.1 features (7 taps, unit-normalised
                          channel differences, 1x1 linear heads, spatial mean) on 3-channel inputs

Random-init weights (the pretrained LPIPS / SqueezeNet weights need the network; SURVEY.md 8d C3 says so): the FLOPs, memory
traffic and launch counts are those of the reference's nets, the values are not -- these nets are outside the parity
contract and nothing in the product imports this file.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class PatchDiscriminatorRef(nn.Module):
    def __init__(self, in_channels=1, channels=32, num_layers_d=3, out_channels=1):
        super().__init__()
        layers = [nn.Sequential(nn.Conv2d(in_channels, channels, 4, 2, 1), nn.LeakyReLU(0.2))]
        cin, cout = channels, channels * 2
        for l_ in range(num_layers_d):
            stride = 2 if l_ != num_layers_d - 1 else 1
            layers.append(nn.Sequential(nn.Conv2d(cin, cout, 4, stride, 1, bias=False), nn.InstanceNorm2d(cout),
                                        nn.LeakyReLU(0.2)))
            cin, cout = cout, cout * 2
        layers.append(nn.Conv2d(cin, out_channels, 4, 1, 1))
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        outs = []
        for layer in self.layers:
            x = layer(x)
            outs.append(x)
        return outs          # the reference reads [-1] (train_vae.py:400,451,453)


def lsgan(logits, target_is_real: bool):
    """PatchAdversarialLoss(criterion="least_squares") for one discriminator output."""
    return F.mse_loss(logits, torch.ones_like(logits) if target_is_real else torch.zeros_like(logits))


class LPIPSSqueezeRef(nn.Module):
    def __init__(self):
        super().__init__()
        import torchvision
        feats = torchvision.models.squeezenet1_1(weights=None).features
        cuts = [2, 5, 8, 10, 11, 12, 13]                     # LPIPS' squeezenet slices
        self.slices = nn.ModuleList()
        prev = 0
        for c in cuts:
            self.slices.append(nn.Sequential(*[feats[i] for i in range(prev, c)]))
            prev = c
        chns = [64, 128, 256, 384, 384, 512, 512]
        self.lins = nn.ModuleList([nn.Conv2d(c, 1, 1, bias=False) for c in chns])
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188]).view(1, 3, 1, 1))
        self.register_buffer("scale", torch.tensor([.458, .448, .450]).view(1, 3, 1, 1))
        for p in self.parameters():
            p.requires_grad_(False)                           # the perceptual net is frozen in the reference too

    def _feats(self, x):
        x = (x - self.shift) / self.scale
        out = []
        for s in self.slices:
            x = s(x)
            out.append(x / (x.pow(2).sum(1, keepdim=True).sqrt() + 1e-10))
        return out

    def forward(self, a, b):
        fa, fb = self._feats(a), self._feats(b)
        val = 0.0
        for x, y, lin in zip(fa, fb, self.lins):
            val = val + lin((x - y) ** 2).mean(dim=(2, 3), keepdim=True)
        return val.mean()


def ensure_three_channels(t):
    return t if t.shape[1] == 3 else t.repeat(1, 3, 1, 1)
