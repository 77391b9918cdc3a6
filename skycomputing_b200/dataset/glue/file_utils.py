"""Local-first file cache (capability parity with scaelum/dataset/glue/file_utils.py:44-259).

The reference module is the HuggingFace/AllenNLP download cache: ``cached_path`` maps an S3/HTTP
URL or a path to a local file, downloading into ``PYTORCH_PRETRAINED_BERT_CACHE`` with an ETag
keyed file name plus a ``.json`` side-car.  Clusters that run this framework are usually offline,
so the order here is: (1) an existing local path wins, (2) a previously cached copy of the URL wins
(any ETag), (3) only then is the network tried (``requests`` for http(s); ``boto3`` for ``s3://``
when installed), and a failure raises ``FileNotFoundError`` naming the cache file a user can drop
in by hand.
"""
from __future__ import annotations

import fnmatch
import hashlib
import json
import os
import shutil
import tempfile
from pathlib import Path
from typing import Optional, Set, Tuple
from urllib.parse import urlparse

PYTORCH_PRETRAINED_BERT_CACHE = Path(os.getenv(
    "PYTORCH_PRETRAINED_BERT_CACHE", Path.home() / ".pytorch_pretrained_bert"))


def _cache_dir(cache_dir) -> str:
    d = str(cache_dir) if cache_dir is not None else str(PYTORCH_PRETRAINED_BERT_CACHE)
    return d


def url_to_filename(url: str, etag: Optional[str] = None) -> str:
    """sha256(url)[.sha256(etag)]: stable, collision-free cache file name."""
    name = hashlib.sha256(url.encode("utf-8")).hexdigest()
    if etag:
        name += "." + hashlib.sha256(etag.encode("utf-8")).hexdigest()
    return name


def filename_to_url(filename: str, cache_dir=None) -> Tuple[str, Optional[str]]:
    """Inverse of :func:`url_to_filename` through the ``<file>.json`` side-car."""
    path = os.path.join(_cache_dir(cache_dir), filename)
    if not os.path.exists(path):
        raise FileNotFoundError(f"file {path} not found")
    meta = path + ".json"
    if not os.path.exists(meta):
        raise FileNotFoundError(f"file {meta} not found")
    with open(meta, encoding="utf-8") as f:
        m = json.load(f)
    return m["url"], m.get("etag")


def cached_path(url_or_filename, cache_dir=None) -> str:
    """Path of a local copy of ``url_or_filename`` (downloads only as a last resort)."""
    s = str(url_or_filename)
    scheme = urlparse(s).scheme
    if scheme in ("http", "https", "s3"):
        return get_from_cache(s, cache_dir)
    if os.path.exists(s):
        return s
    if scheme == "":
        raise FileNotFoundError(f"file {s} not found")
    raise ValueError(f"unable to parse {s} as a URL or as a local path")


def split_s3_path(url: str) -> Tuple[str, str]:
    parsed = urlparse(url)
    if not parsed.netloc or not parsed.path:
        raise ValueError(f"bad s3 path {url}")
    key = parsed.path[1:] if parsed.path.startswith("/") else parsed.path
    return parsed.netloc, key


def s3_request(func):
    """Decorator: turn a missing-object error of an S3 call into FileNotFoundError."""
    def wrapper(url, *args, **kwargs):
        try:
            return func(url, *args, **kwargs)
        except Exception as exc:  # botocore.exceptions.ClientError when boto3 is present
            code = getattr(exc, "response", {}).get("Error", {}).get("Code") if hasattr(exc, "response") else None
            if code is not None and int(code) == 404:
                raise FileNotFoundError(f"file {url} not found") from exc
            raise
    wrapper.__name__ = getattr(func, "__name__", "s3_call")
    return wrapper


def _s3_resource():
    try:
        import boto3
    except ImportError as exc:
        raise FileNotFoundError("s3:// URLs need boto3, which is not installed") from exc
    return boto3.resource("s3")


@s3_request
def s3_etag(url: str) -> Optional[str]:
    bucket, key = split_s3_path(url)
    return _s3_resource().Object(bucket, key).e_tag


@s3_request
def s3_get(url: str, temp_file) -> None:
    bucket, key = split_s3_path(url)
    _s3_resource().Bucket(bucket).download_fileobj(key, temp_file)


def http_get(url: str, temp_file) -> None:
    import requests

    with requests.get(url, stream=True, timeout=30) as req:
        req.raise_for_status()
        for chunk in req.iter_content(chunk_size=1 << 16):
            if chunk:
                temp_file.write(chunk)


def _cached_copy(url: str, cache_dir: str) -> Optional[str]:
    """Any cached file of this URL, whatever ETag it was fetched under (newest first)."""
    stem = url_to_filename(url)
    if not os.path.isdir(cache_dir):
        return None
    hits = [f for f in os.listdir(cache_dir)
            if (f == stem or fnmatch.fnmatch(f, stem + ".*")) and not f.endswith(".json")]
    if not hits:
        return None
    hits.sort(key=lambda f: os.path.getmtime(os.path.join(cache_dir, f)), reverse=True)
    return os.path.join(cache_dir, hits[0])


def get_from_cache(url: str, cache_dir=None) -> str:
    cache_dir = _cache_dir(cache_dir)
    os.makedirs(cache_dir, exist_ok=True)
    hit = _cached_copy(url, cache_dir)
    if hit is not None:
        return hit
    etag = None
    try:
        if url.startswith("s3://"):
            etag = s3_etag(url)
        else:
            import requests

            head = requests.head(url, allow_redirects=True, timeout=10)
            if head.status_code != 200:
                raise IOError(f"HEAD request failed for url {url} with status code {head.status_code}")
            etag = head.headers.get("ETag")
    except FileNotFoundError:
        raise
    except Exception as exc:
        raise FileNotFoundError(
            f"{url} is not cached and cannot be fetched ({exc}); place the file at "
            f"{os.path.join(cache_dir, url_to_filename(url))}") from exc
    path = os.path.join(cache_dir, url_to_filename(url, etag))
    if not os.path.exists(path):
        with tempfile.NamedTemporaryFile() as tmp:
            if url.startswith("s3://"):
                s3_get(url, tmp)
            else:
                http_get(url, tmp)
            tmp.flush()
            tmp.seek(0)
            with open(path, "wb") as out:
                shutil.copyfileobj(tmp, out)
        with open(path + ".json", "w", encoding="utf-8") as f:
            json.dump({"url": url, "etag": etag}, f)
    return path


def read_set_from_file(filename: str) -> Set[str]:
    """One item per line -> set (trailing whitespace stripped)."""
    with open(filename, encoding="utf-8") as f:
        return {line.rstrip() for line in f}


def get_file_extension(path: str, dot: bool = True, lower: bool = True) -> str:
    ext = os.path.splitext(path)[1]
    ext = ext if dot else ext[1:]
    return ext.lower() if lower else ext
