from .yolo import YOLO
from .reid import ReID
from .label import set_label_map, get_label_name
