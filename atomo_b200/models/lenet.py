"""LeNet (parity: ``/root/reference/src/model_ops/lenet.py:12-35``).

conv 1->20 k5, conv 20->50 k5, fc 800->500, fc 500->10; max-pool *before* ReLU
and no activation between fc1 and fc2, exactly as the reference wires it, so
parameter shapes (and therefore the matricized gradient shapes of SURVEY.md
2.4) match.
"""
import torch.nn as nn
import torch.nn.functional as F


class LeNet(nn.Module):
    def __init__(self, num_classes: int = 10):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 20, 5, 1)
        self.conv2 = nn.Conv2d(20, 50, 5, 1)
        self.fc1 = nn.Linear(4 * 4 * 50, 500)
        self.fc2 = nn.Linear(500, num_classes)
        self.full_modules = [self.conv1, self.conv2, self.fc1, self.fc2]

    def forward(self, x):
        x = F.relu(F.max_pool2d(self.conv1(x), 2, 2))
        x = F.relu(F.max_pool2d(self.conv2(x), 2, 2))
        x = x.reshape(-1, 4 * 4 * 50)
        return self.fc2(self.fc1(x))

    def name(self):
        return "lenet"
