"""Class-label names (role of fastmot/models/label.py:104-121: `set_label_map`, `get_label_name`)."""

_COCO = (
    "person bicycle car motorcycle airplane bus train truck boat traffic_light fire_hydrant stop_sign "
    "parking_meter bench bird cat dog horse sheep cow elephant bear zebra giraffe backpack umbrella handbag tie "
    "suitcase frisbee skis snowboard sports_ball kite baseball_bat baseball_glove skateboard surfboard "
    "tennis_racket bottle wine_glass cup fork knife spoon bowl banana apple sandwich orange broccoli carrot "
    "hot_dog pizza donut cake chair couch potted_plant bed dining_table toilet tv laptop mouse remote keyboard "
    "cell_phone microwave oven toaster sink refrigerator book clock vase scissors teddy_bear hair_drier "
    "toothbrush").split()

LABEL_MAP = list(_COCO)


def set_label_map(label_map):
    """Set label name mapping from class IDs (app.py:64)."""
    global LABEL_MAP
    LABEL_MAP = list(label_map)


def get_label_name(class_id):
    class_id = int(class_id)
    if 0 <= class_id < len(LABEL_MAP):
        return LABEL_MAP[class_id]
    return f"class{class_id}"
