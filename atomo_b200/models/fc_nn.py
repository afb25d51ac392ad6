"""FC_NN: 784-800-500-10 MLP with a sigmoid output
(parity: ``/root/reference/src/model_ops/fc_nn.py:12-30``)."""
import torch.nn as nn


class FC_NN(nn.Module):
    def __init__(self, num_classes: int = 10):
        super().__init__()
        self.fc1 = nn.Linear(784, 800)
        self.fc2 = nn.Linear(800, 500)
        self.fc3 = nn.Linear(500, num_classes)
        self.relu = nn.ReLU()
        self.sigmoid = nn.Sigmoid()
        self.full_modules = [self.fc1, self.fc2, self.fc3]

    def forward(self, x):
        x = x.reshape(x.size(0), -1)
        x = self.relu(self.fc1(x))
        x = self.relu(self.fc2(x))
        return self.sigmoid(self.fc3(x))

    def name(self):
        return "fc_nn"
