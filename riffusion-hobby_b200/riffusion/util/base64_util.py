"""base64 helper of the response contract (reference: riffusion/util/base64_util.py)."""
import base64
import io


def encode(buffer: io.BytesIO) -> str:
    """bytes of the buffer as base64 text with the newline-wrapped framing `base64.encodebytes` produces"""
    return base64.encodebytes(buffer.getvalue()).decode("ascii")
