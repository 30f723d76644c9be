class PpoPlayerContinuous:
    def __init__(self, config): self.config = config
