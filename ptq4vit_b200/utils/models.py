"""Minimal ViT / DeiT definitions with timm's submodule names (patch_embed.proj, blocks.N.attn.{qkv,proj},
blocks.N.mlp.{fc1,fc2}, head) and the reference's attention rewrite that routes q@k^T and attn@v through
`matmul1` / `matmul2` modules (utils/models.py:10-26, :58-60).  timm and pretrained weights are not
available offline, so weights are synthetic (trunc-normal 0.02) -- this is harness code, not the hot path."""
import torch
import torch.nn as nn


class MatMul(nn.Module):
    def forward(self, A, B):
        return A @ B


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)
        self.matmul1 = MatMul()
        self.matmul2 = MatMul()

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = self.matmul1(q, k.transpose(-2, -1)) * self.scale
        attn = attn.softmax(dim=-1)
        x = self.matmul2(attn, v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch, dim):
        super().__init__()
        self.num_patches = (img_size // patch) ** 2
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch=16, dim=768, depth=12, num_heads=12, num_classes=1000, seed=0):
        super().__init__()
        self.patch_embed = PatchEmbed(img_size, patch, dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, dim))
        self.blocks = nn.Sequential(*[Block(dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = nn.Linear(dim, num_classes)
        gen = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for name_, p_ in self.named_parameters():
                if p_.dim() > 1:
                    p_.copy_(torch.nn.init.trunc_normal_(torch.empty_like(p_), std=0.02, generator=gen))
                elif name_.endswith("bias") and "norm" not in name_:
                    # nn.Linear / nn.Conv2d draw their biases from the global RNG: make the whole net a function of `seed`
                    p_.copy_(torch.empty_like(p_).uniform_(-0.02, 0.02, generator=gen))
            # synthetic nets have no trained structure: widen activations so that the blocks differ
            for blk in self.blocks:
                blk.attn.qkv.weight.mul_(4.0)
                blk.mlp.fc1.weight.mul_(4.0)
            self.head.weight.mul_(8.0)
            self.pos_embed.copy_(torch.nn.init.trunc_normal_(torch.empty_like(self.pos_embed), std=0.02, generator=gen))

    def forward(self, x):
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.pos_embed
        x = self.norm(self.blocks(x))
        return self.head(x[:, 0])


_ZOO = {
    "vit_tiny_patch16_224": dict(img_size=224, patch=16, dim=192, depth=12, num_heads=3),
    "vit_small_patch16_224": dict(img_size=224, patch=16, dim=384, depth=12, num_heads=6),
    "vit_base_patch16_224": dict(img_size=224, patch=16, dim=768, depth=12, num_heads=12),
    "vit_base_patch16_384": dict(img_size=384, patch=16, dim=768, depth=12, num_heads=12),
    "deit_small_patch16_224": dict(img_size=224, patch=16, dim=384, depth=12, num_heads=6),
    "deit_base_patch16_224": dict(img_size=224, patch=16, dim=768, depth=12, num_heads=12),
    "deit_base_patch16_384": dict(img_size=384, patch=16, dim=768, depth=12, num_heads=12),
}


def get_net(name, device="cuda", seed=0, **override):
    """reference: utils/models.py:62-91 (timm.create_model + attention rewrite); here: synthetic weights."""
    cfg = dict(_ZOO[name]); cfg.update(override)
    net = VisionTransformer(seed=seed, **cfg)
    return net.to(device).eval()
