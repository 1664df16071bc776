"""JSON decoder that turns arrays into tuples (fastmot/utils/decoder.py:4-14), so cfg/mot.json values can
be splatted into constructors exactly as app.py:57-58 does."""
import json


class ConfigDecoder(json.JSONDecoder):
    def __init__(self, **kwargs):
        json.JSONDecoder.__init__(self, **kwargs)
        self.parse_array = self._parse_array
        self.scan_once = json.scanner.py_make_scanner(self)

    def _parse_array(self, *args, **kwargs):
        values, end = json.decoder.JSONArray(*args, **kwargs)
        return tuple(values), end
