class DefaultRewardsShaper:
    def __init__(self, scale_value=1, shift_value=0, min_val=-float('inf'), max_val=float('inf'), is_torch=True):
        self.scale_value = scale_value
        self.shift_value = shift_value
    def __call__(self, reward):
        reward = reward + self.shift_value
        reward = reward * self.scale_value
        return reward
