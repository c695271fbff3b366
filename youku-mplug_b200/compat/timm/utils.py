def get_state_dict(model, unwrap_fn=None):
    return model.state_dict()


class ModelEma:
    def __init__(self, *a, **k):
        raise NotImplementedError("timm.utils.ModelEma is not provided by the stand-in")
