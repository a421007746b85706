"""CSV -> MARO ``.bin`` converter, byte-compatible with the reference's (maro/data_lib/binary_converter.py:68-228,
item_meta.py:45-285, common.py:7-44).

File layout (little endian):
    header   struct "<4s b I Q I QQ QQ qq": b"MARO", file type 1, version 100, item count, item size, meta offset, meta size,
             data offset, data size, start time, end time (UTC seconds of the smallest / largest ``timestamp``)
    meta     YAML text: attributes (``!MaroAttribute`` mappings, ``timestamp`` first), events (``!MaroEvent``),
             default_event_name, event_attr_name
    items    one packed struct per CSV row, fields in attribute order

Behaviour that has to be matched to stay byte-identical (each checked against the reference's own csv/bin fixture pairs):
  * the ``timestamp`` attribute moves to the front but keeps its configured dtype (item_meta.py:248-254);
  * the per-attribute ``tzone`` is dropped when the attribute is built (item_meta.py:33) and the meta never gets a time zone,
    so datetime strings are read as UTC (binary_converter.py:15-31, 56-61);
  * numbers go through ``float`` first (``int(float("1.9")) == 1``), quotes and blanks around a value are stripped;
  * a row with an unparsable value is skipped with a warning, a column missing from the CSV packs as 0.
"""
from __future__ import annotations

import calendar
import csv
import re
import struct
import warnings
from typing import Dict, List, Optional

import yaml

_HEADER = struct.Struct("<4s b I Q I QQ QQ qq")
_VERSION, _SINGLE_FILE = 100, 1
_PACK = {"i": "i", "i4": "i", "i2": "h", "i8": "q", "f": "f", "d": "d"}
_PY = {"i": int, "i2": int, "i4": int, "i8": int, "f": float, "d": float}


class _Attribute:
    """one ``!MaroAttribute`` of the meta block (field names are the file format's)"""

    def __init__(self, name, dtype, slot, raw_name, adjust_ratio):
        self.name, self.dtype, self.slot, self.raw_name, self.adjust_ratio, self.tzone = name, dtype, slot, raw_name, adjust_ratio, None


class _Event:
    """one ``!MaroEvent`` of the meta block"""

    def __init__(self, display_name, type_name, value):
        self.display_name, self.type_name, self.value = display_name, type_name, value


class _MetaDumper(yaml.SafeDumper):
    pass


_MetaDumper.add_representer(_Attribute, lambda d, o: d.represent_mapping("!MaroAttribute", o.__dict__))
_MetaDumper.add_representer(_Event, lambda d, o: d.represent_mapping("!MaroEvent", o.__dict__))


class BinaryMeta:
    """The validated meta of one binary file (item_meta.py:45)."""

    def __init__(self, conf: dict):
        self.attributes: List[_Attribute] = []
        self.events: List[_Event] = []
        entity = conf.get("entity", {}) or {}
        self.event_attr_name = entity.get("_event", None)
        has_timestamp = False
        for name, settings in entity.items():
            if type(settings) != dict:
                continue
            dtype = settings.get("dtype", "i")
            if dtype in _PACK and re.match(r"^[a-z A-Z]+", name):
                attr = _Attribute(name, dtype, settings.get("slot", 1), settings.get("column", None), settings.get("adjust_ratio"))
                if name == "timestamp":
                    has_timestamp = True
                    self.attributes.insert(0, attr)
                else:
                    self.attributes.append(attr)
            else:
                warnings.warn(f"invalid attribute {name}, ignore it")
        if not has_timestamp:
            raise ValueError("the meta must define a 'timestamp' attribute (MetaTimestampNotExist in the reference)")
        events = conf.get("events", {}) or {}
        self.default_event_name = events.get("_default", None)
        for type_name, settings in events.items():
            if type(settings) != dict:
                continue
            self.events.append(_Event(settings.get("display_name", type_name), type_name, settings.get("value_in_csv", None)))
        self.item_struct = struct.Struct("<" + "".join(_PACK[a.dtype] for a in self.attributes))

    @classmethod
    def from_file(cls, path: str) -> "BinaryMeta":
        with open(path, "rt") as fp:
            return cls(yaml.safe_load(fp))

    @property
    def item_size(self) -> int:
        return self.item_struct.size

    def to_bytes(self) -> bytes:
        return yaml.dump({"events": self.events, "attributes": self.attributes, "default_event_name": self.default_event_name,
                          "event_attr_name": self.event_attr_name}, Dumper=_MetaDumper).encode()


def _parse_value(text: str, dtype: str):
    """binary_converter.py:34-65: number via float(), else a datetime read as UTC -> epoch seconds; None if neither"""
    text = text.strip("\"'").strip()
    try:
        return _PY[dtype](float(text))
    except ValueError:
        pass
    try:
        from dateutil.parser import parse as parse_dt
        from dateutil.tz import UTC

        dt = parse_dt(text).replace(tzinfo=UTC)
        return calendar.timegm(dt.astimezone(UTC).timetuple())
    except Exception:
        warnings.warn(f"Cannot parse value '{text}' into type '{dtype}'")
        return None


class BinaryConverter:
    """``BinaryConverter(output_file, meta_file[, utc_start_timestamp])``; ``add_csv(path)`` any number of times (rows are
    appended in call order, not sorted); ``flush()`` rewrites the header.  Closing happens on ``flush`` / ``close`` / GC."""

    def __init__(self, output_file: str, meta_file: str, utc_start_timestamp: Optional[int] = None):
        self._meta = BinaryMeta.from_file(meta_file)
        self._fp = open(output_file, "wb+")
        self._item_count = self._data_size = self._endtime = 0
        self._starttime = 0 if utc_start_timestamp is None else utc_start_timestamp
        self._start_fixed = utc_start_timestamp is not None
        meta_bytes = self._meta.to_bytes()
        self._meta_offset, self._meta_size = _HEADER.size, len(meta_bytes)
        self._data_offset = self._meta_offset + self._meta_size
        self._write_header()
        self._fp.write(meta_bytes)

    @property
    def meta(self) -> BinaryMeta:
        return self._meta

    def _write_header(self):
        self._fp.seek(0, 0)
        self._fp.write(_HEADER.pack(b"MARO", _SINGLE_FILE, _VERSION, self._item_count, self._meta.item_size, self._meta_offset,
                                    self._meta_size, self._data_offset, self._data_size, self._starttime, self._endtime))
        self._fp.seek(0, 2)

    def add_csv(self, csv_file: str):
        attrs = self._meta.attributes
        pack = self._meta.item_struct.pack
        with open(csv_file, newline="") as fp:
            for row in csv.DictReader(fp):
                values, ok = [0] * len(attrs), True
                for k, a in enumerate(attrs):
                    if a.raw_name not in row:  # a field the CSV does not have packs as zero
                        continue
                    v = _parse_value(row[a.raw_name], a.dtype)
                    if v is None:
                        ok = False
                        break
                    values[k] = v
                    if a.name == "timestamp":
                        self._starttime = v if not self._start_fixed else min(self._starttime, v)
                        self._start_fixed = True
                        self._endtime = max(v, self._endtime)
                if ok:
                    self._fp.write(pack(*values))
                    self._item_count += 1
                    self._data_size += self._meta.item_size

    def flush(self):
        if self._fp is not None and not self._fp.closed:
            self._write_header()
            self._fp.flush()

    def close(self):
        if self._fp is not None and not self._fp.closed:
            self.flush()
            self._fp.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
