"""Stand-in for `mlx.optimizers` (TEST INFRASTRUCTURE ONLY): only the constructor is needed so that
`create_NeRF` from the reference can be executed; the update rule is restated in oracle/training.py."""


class Adam:
    def __init__(self, learning_rate, betas=(0.9, 0.999), eps=1e-8):
        self.learning_rate = learning_rate
        self.betas = betas
        self.eps = eps
        self.state = {}
