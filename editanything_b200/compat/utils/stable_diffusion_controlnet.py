from editanything_b200.loading import ControlNetModel, ControlNetModel2  # noqa: F401
