"""keras.optimizers (2.0.8): Adam and RMSprop update rules."""
import math

import torch


class Optimizer:
    def __init__(self, lr):
        self.lr, self.iterations, self.state = float(lr), 0, {}


class Adam(Optimizer):
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-8, decay=0.0):
        super().__init__(lr)
        self.beta_1, self.beta_2, self.epsilon, self.decay = beta_1, beta_2, epsilon, decay

    def apply(self, params, grads):
        lr = self.lr
        if self.decay > 0:
            lr *= 1.0 / (1.0 + self.decay * self.iterations)
        t = self.iterations + 1
        lr_t = lr * (math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t))
        with torch.no_grad():
            for p, g in zip(params, grads):
                m, v = self.state.setdefault(id(p), (torch.zeros_like(p), torch.zeros_like(p)))
                m.mul_(self.beta_1).add_((1.0 - self.beta_1) * g)
                v.mul_(self.beta_2).add_((1.0 - self.beta_2) * g * g)
                p.sub_(lr_t * m / (torch.sqrt(v) + self.epsilon))
        self.iterations += 1


class RMSprop(Optimizer):
    def __init__(self, lr=0.001, rho=0.9, epsilon=1e-8, decay=0.0):
        super().__init__(lr)
        self.rho, self.epsilon = rho, epsilon

    def apply(self, params, grads):
        with torch.no_grad():
            for p, g in zip(params, grads):
                a = self.state.setdefault(id(p), torch.zeros_like(p))
                a.mul_(self.rho).add_((1.0 - self.rho) * g * g)
                p.sub_(self.lr * g / (torch.sqrt(a) + self.epsilon))
        self.iterations += 1
