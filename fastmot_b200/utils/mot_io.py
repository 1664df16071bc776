"""MOT Challenge result lines, as written by the reference application (app.py:91-97).

One line per visible track and frame: `frame,id,left,top,width,height,-1,-1,-1` with the box mapped from the
processing resolution back to the stream resolution and `width = right - left + 1`.
"""
import numpy as np


def mot_result_line(frame_count, trk_id, tlbr, resize_to, resolution):
    tlbr = np.asarray(tlbr, np.float64)
    resize_to = np.asarray(resize_to, np.float64)
    resolution = np.asarray(resolution, np.float64)
    tl = tlbr[:2] / resize_to * resolution
    br = tlbr[2:] / resize_to * resolution
    w, h = br - tl + 1
    return f'{frame_count},{trk_id},{tl[0]:.6f},{tl[1]:.6f},{w:.6f},{h:.6f},-1,-1,-1\n'


def write_mot_results(txt, mot, resize_to, resolution):
    """Append the current frame's visible tracks of a `MOT` instance to the open text file `txt`."""
    for track in mot.visible_tracks():
        txt.write(mot_result_line(mot.frame_count, track.trk_id, track.tlbr, resize_to, resolution))
