"""ReID model descriptors (API of fastmot/models/reid.py:10-45, values :95-109)."""


class ReID:
    __registry = {}

    ARCH = None          # ('osnet', width multiplier): built-in architecture, synthetic weights
    MODEL_PATH = None    # path to an ONNX file (reid.py:20-23); imported by models/onnx_import.py when set
    WEIGHTS_PATH = None
    INPUT_SHAPE = None
    OUTPUT_LAYOUT = None
    METRIC = None

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        cls.__registry[cls.__name__] = cls

    @classmethod
    def get_model(cls, name):
        return cls.__registry[name]


class OSNet025(ReID):
    ARCH = ('osnet', 0.25)
    INPUT_SHAPE = (3, 256, 128)
    OUTPUT_LAYOUT = 512
    METRIC = 'euclidean'


class OSNet10(ReID):
    ARCH = ('osnet', 1.0)
    INPUT_SHAPE = (3, 256, 128)
    OUTPUT_LAYOUT = 512
    METRIC = 'cosine'
