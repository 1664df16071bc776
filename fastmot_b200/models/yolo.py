"""YOLO model descriptors: same class-attribute "plugin" API and auto-registry as fastmot/models/yolo.py:11-58
(subclass -> registered by name -> selected by the `model` string of the config).  TensorRT engine paths are
replaced by a Darknet-cfg graph description consumed by fastmot_b200.engine (no TensorRT here).

Descriptor values (NUM_CLASSES, LETTERBOX, NEW_COORDS, INPUT_SHAPE, LAYER_FACTORS, SCALES, ANCHORS) follow
fastmot/models/yolo.py:154-299.
"""


class YOLO:
    __registry = {}

    CFG = None            # callable returning Darknet cfg text (fastmot_b200/models/darknet_cfgs.py)
    WEIGHTS_PATH = None   # optional Darknet .weights; None -> seeded synthetic weights
    NUM_CLASSES = None
    LETTERBOX = False
    NEW_COORDS = False
    INPUT_SHAPE = None
    LAYER_FACTORS = None
    SCALES = None
    ANCHORS = None

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        cls.__registry[cls.__name__] = cls

    @classmethod
    def get_model(cls, name):
        return cls.__registry[name]


class YOLOv4(YOLO):
    CFG = 'yolov4'
    NUM_CLASSES = 2
    INPUT_SHAPE = (3, 512, 512)
    LAYER_FACTORS = [8, 16, 32]
    SCALES = [1.2, 1.1, 1.05]
    ANCHORS = [[11, 22, 24, 60, 37, 116],
               [54, 186, 69, 268, 89, 369],
               [126, 491, 194, 314, 278, 520]]


class YOLOv4CSP(YOLO):
    CFG = 'yolov4-csp'
    NUM_CLASSES = 1
    LETTERBOX = True
    NEW_COORDS = True
    INPUT_SHAPE = (3, 640, 640)
    LAYER_FACTORS = [8, 16, 32]
    SCALES = [2.0, 2.0, 2.0]
    ANCHORS = [[12, 16, 19, 36, 40, 28],
               [36, 75, 76, 55, 72, 146],
               [142, 110, 192, 243, 459, 401]]


class YOLOv4P5(YOLO):
    CFG = 'yolov4-p5'
    NUM_CLASSES = 1
    LETTERBOX = True
    NEW_COORDS = True
    INPUT_SHAPE = (3, 896, 896)
    LAYER_FACTORS = [8, 16, 32]
    SCALES = [2.0, 2.0, 2.0]
    ANCHORS = [[13, 17, 31, 25, 24, 51, 61, 45],
               [48, 102, 119, 96, 97, 189, 217, 184],
               [171, 384, 324, 451, 616, 618, 800, 800]]


class YOLOv4P5_1280(YOLOv4P5):
    """BASELINE.json config 4 quotes 'YOLOv4-p5 1280x'; same graph at a 1280 input."""
    INPUT_SHAPE = (3, 1280, 1280)


class YOLOv4Tiny(YOLO):
    CFG = 'yolov4-tiny'
    NUM_CLASSES = 1
    INPUT_SHAPE = (3, 416, 416)
    LAYER_FACTORS = [32, 16]
    SCALES = [1.05, 1.05]
    ANCHORS = [[81, 82, 135, 169, 344, 319],
               [23, 27, 37, 58, 81, 82]]
