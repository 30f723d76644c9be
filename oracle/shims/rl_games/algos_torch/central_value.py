class CentralValueTrain:  # unused: has_central_value is False for every reference config
    def __init__(self, **kw): raise NotImplementedError
