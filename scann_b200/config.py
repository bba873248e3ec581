"""Minimal protobuf text-format reader/writer for ScannConfig-shaped messages.

The reference hands the builder's text proto to `TextFormat::ParseFromString`
(scann_ops/cc/scann.h:185-188).  `scann_builder.py` emits Python literals (`True`, `False`,
`nan`) and both `field { ... }` and `field {k: "v"}` spellings, so the reader accepts those.
Messages are parsed into `Msg` objects: ordered multi-maps of field name -> scalar string / Msg.
"""
import math


class Msg:
  """One parsed message: list of (name, value) with value a str (scalar) or a Msg."""

  def __init__(self):
    self.fields = []

  def add(self, name, value):
    self.fields.append((name, value))

  def all(self, name):
    return [v for n, v in self.fields if n == name]

  def get(self, name, default=None):
    for n, v in self.fields:
      if n == name:
        return v
    return default

  def has(self, name):
    return any(n == name for n, _ in self.fields)

  def path(self, *names, default=None):
    cur = self
    for n in names:
      if not isinstance(cur, Msg):
        return default
      cur = cur.get(n)
      if cur is None:
        return default
    return cur

  def __repr__(self):
    return f"Msg({self.fields!r})"


class TextProtoError(ValueError):
  pass


def parse(text):
  pos = 0
  n = len(text)

  def skip():
    nonlocal pos
    while pos < n:
      c = text[pos]
      if c.isspace() or c in ",;":
        pos += 1
      elif c == "#":
        while pos < n and text[pos] != "\n":
          pos += 1
      else:
        break

  def fields(closer):
    nonlocal pos
    m = Msg()
    while True:
      skip()
      if pos >= n:
        if closer:
          raise TextProtoError("unexpected end of text proto")
        return m
      if closer and text[pos] == closer:
        pos += 1
        return m
      b = pos
      while pos < n and (text[pos].isalnum() or text[pos] in "_."):
        pos += 1
      if b == pos:
        raise TextProtoError(f"unexpected character {text[pos]!r} at offset {pos}")
      name = text[b:pos]
      skip()
      colon = False
      if pos < n and text[pos] == ":":
        colon = True
        pos += 1
        skip()
      if pos < n and text[pos] in "{<":
        close = "}" if text[pos] == "{" else ">"
        pos += 1
        m.add(name, fields(close))
        continue
      if not colon:
        raise TextProtoError(f"expected ':' or '{{' after field {name}")
      if pos < n and text[pos] in "\"'":
        q = text[pos]
        pos += 1
        out = []
        while pos < n and text[pos] != q:
          if text[pos] == "\\" and pos + 1 < n:
            pos += 1
            out.append({"n": "\n", "t": "\t"}.get(text[pos], text[pos]))
          else:
            out.append(text[pos])
          pos += 1
        if pos >= n:
          raise TextProtoError("unterminated string")
        pos += 1
        m.add(name, "".join(out))
      else:
        b = pos
        while pos < n and not text[pos].isspace() and text[pos] not in ",;}>#":
          pos += 1
        if b == pos:
          raise TextProtoError(f"missing value for field {name}")
        m.add(name, text[b:pos])

  return fields(None)


def as_bool(v, default=False):
  if v is None:
    return default
  s = str(v)
  if s in ("true", "True", "t", "1"):
    return True
  if s in ("false", "False", "f", "0"):
    return False
  raise TextProtoError(f"bad bool {v!r}")


def as_float(v, default=None):
  if v is None:
    return default
  s = str(v).lower()
  if s == "nan":
    return math.nan
  if s in ("inf", "+inf", "infinity"):
    return math.inf
  if s in ("-inf", "-infinity"):
    return -math.inf
  return float(s.rstrip("f"))


def as_int(v, default=None):
  if v is None:
    return default
  return int(str(v), 0)


# ---- writer ------------------------------------------------------------------------------------
class Enum(str):
  """A bare (unquoted) enum identifier in the emitted text."""


def emit(msg, indent=0):
  """msg: list of (name, value); value is a list (sub-message), Enum, str, bool, int or float."""
  pad = "  " * indent
  lines = []
  for name, value in msg:
    if value is None:
      continue
    if isinstance(value, list):
      lines.append(f"{pad}{name} {{")
      lines.append(emit(value, indent + 1))
      lines.append(f"{pad}}}")
    elif isinstance(value, Enum):
      lines.append(f"{pad}{name}: {value}")
    elif isinstance(value, bool):
      lines.append(f"{pad}{name}: {'true' if value else 'false'}")
    elif isinstance(value, str):
      esc = value.replace("\\", "\\\\").replace('"', '\\"')
      lines.append(f'{pad}{name}: "{esc}"')
    elif isinstance(value, float):
      if math.isnan(value):
        lines.append(f"{pad}{name}: nan")
      elif math.isinf(value):
        lines.append(f"{pad}{name}: {'inf' if value > 0 else '-inf'}")
      else:
        lines.append(f"{pad}{name}: {value!r}")
    else:
      lines.append(f"{pad}{name}: {value}")
  return "\n".join(l for l in lines if l != "")
