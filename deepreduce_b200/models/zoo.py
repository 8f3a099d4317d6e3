"""Remaining model families of the reference's experiments (paper Table 1, Tables
2/5/6; reference README.md:20-22, tensorflow/deepreduce.py:182-253 names
resnet20_v2 / vgg16 / resnet50): VGG-16, DenseNet40-K12, MobileNet, NCF (NeuMF,
MovieLens-20M shapes), the StackOverflow next-word LSTM, and BERT-large
(BASELINE.json config 5).  All random-init, synthetic-data friendly.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---- VGG-16 (CIFAR/ImageNet) -------------------------------------------------
class VGG16(nn.Module):
    cfg = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']

    def __init__(self, num_classes: int = 10, in_hw: int = 32):
        super().__init__()
        layers, c = [], 3
        for v in self.cfg:
            if v == 'M':
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(c, v, 3, padding=1), nn.BatchNorm2d(v), nn.ReLU(inplace=True)]
                c = v
        self.features = nn.Sequential(*layers)
        hw = in_hw // 32
        self.classifier = nn.Sequential(nn.Linear(512 * hw * hw, 512), nn.ReLU(inplace=True), nn.Linear(512, num_classes))

    def forward(self, x):
        return self.classifier(self.features(x).flatten(1))


# ---- DenseNet40-K12 ----------------------------------------------------------
class _DenseLayer(nn.Module):
    def __init__(self, inp, growth):
        super().__init__()
        self.bn = nn.BatchNorm2d(inp)
        self.conv = nn.Conv2d(inp, growth, 3, padding=1, bias=False)

    def forward(self, x):
        return torch.cat([x, self.conv(F.relu(self.bn(x)))], 1)


class DenseNet40(nn.Module):
    def __init__(self, growth: int = 12, num_classes: int = 10):
        super().__init__()
        n = (40 - 4) // 3
        c = 2 * growth
        self.conv1 = nn.Conv2d(3, c, 3, padding=1, bias=False)
        blocks = []
        for b in range(3):
            for _ in range(n):
                blocks.append(_DenseLayer(c, growth))
                c += growth
            if b < 2:
                blocks += [nn.BatchNorm2d(c), nn.ReLU(inplace=True), nn.Conv2d(c, c, 1, bias=False), nn.AvgPool2d(2)]
        self.blocks = nn.Sequential(*blocks)
        self.bn = nn.BatchNorm2d(c)
        self.fc = nn.Linear(c, num_classes)

    def forward(self, x):
        x = self.blocks(self.conv1(x))
        x = F.adaptive_avg_pool2d(F.relu(self.bn(x)), 1).flatten(1)
        return self.fc(x)


# ---- MobileNet v1 (CIFAR variant) ---------------------------------------------
class MobileNet(nn.Module):
    cfg = [64, (128, 2), 128, (256, 2), 256, (512, 2), 512, 512, 512, 512, 512, (1024, 2), 1024]

    def __init__(self, num_classes: int = 10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 32, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        layers, c = [], 32
        for v in self.cfg:
            out, s = (v, 1) if isinstance(v, int) else v
            layers += [nn.Conv2d(c, c, 3, s, 1, groups=c, bias=False), nn.BatchNorm2d(c), nn.ReLU(inplace=True),
                       nn.Conv2d(c, out, 1, bias=False), nn.BatchNorm2d(out), nn.ReLU(inplace=True)]
            c = out
        self.layers = nn.Sequential(*layers)
        self.fc = nn.Linear(1024, num_classes)

    def forward(self, x):
        x = self.layers(F.relu(self.bn1(self.conv1(x))))
        return self.fc(F.adaptive_avg_pool2d(x, 1).flatten(1))


# ---- NCF / NeuMF (MovieLens-20M: 138 493 users x 26 744 items -> 31 832 577 params) ----
class NeuMF(nn.Module):
    def __init__(self, n_users: int = 138493, n_items: int = 26744, mf_dim: int = 64,
                 mlp_layers=(256, 256, 128, 64)):
        super().__init__()
        self.mf_user = nn.Embedding(n_users, mf_dim)
        self.mf_item = nn.Embedding(n_items, mf_dim)
        self.mlp_user = nn.Embedding(n_users, mlp_layers[0] // 2)
        self.mlp_item = nn.Embedding(n_items, mlp_layers[0] // 2)
        self.mlp = nn.ModuleList(nn.Linear(a, b) for a, b in zip(mlp_layers[:-1], mlp_layers[1:]))
        self.out = nn.Linear(mf_dim + mlp_layers[-1], 1)
        for e in (self.mf_user, self.mf_item, self.mlp_user, self.mlp_item):
            nn.init.normal_(e.weight, 0.0, 0.01)

    def forward(self, user, item):
        mf = self.mf_user(user) * self.mf_item(item)
        x = torch.cat([self.mlp_user(user), self.mlp_item(item)], dim=1)
        for l in self.mlp:
            x = F.relu(l(x))
        return self.out(torch.cat([mf, x], dim=1)).squeeze(-1)


# ---- StackOverflow next-word LSTM (paper Table 1: 4 053 428 params) ------------
class NextWordLSTM(nn.Module):
    def __init__(self, vocab: int = 10004, embed: int = 96, hidden: int = 670, layers: int = 1):
        super().__init__()
        self.embed = nn.Embedding(vocab, embed)
        self.lstm = nn.LSTM(embed, hidden, num_layers=layers, batch_first=True)
        self.proj = nn.Linear(hidden, embed)
        self.out = nn.Linear(embed, vocab)

    def forward(self, tokens):
        h, _ = self.lstm(self.embed(tokens))
        return self.out(self.proj(h))


# ---- BERT-large (random init through transformers' config) ---------------------
def bert_large(vocab_size: int = 30522, seq_len: int = 512):
    from transformers import BertConfig, BertForMaskedLM
    cfg = BertConfig(vocab_size=vocab_size, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                     intermediate_size=4096, max_position_embeddings=seq_len)
    return BertForMaskedLM(cfg)


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
