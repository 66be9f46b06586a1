"""Pins oracle/gunzip_oracle.py (the restatement of Go's compress/gzip reader that cmd/fqpack/main.go:142-174 puts in
front of compress.Compress): against the reference's own gzip cases (cmd/fqpack/main_test.go: gzip by extension, gzip by
magic, plain input left alone) and against Python's independent gzip module on every generated file."""
import gzip
import io

import pytest

from oracle import gunzip_oracle as go
from tests.gzip_cases import bad_cases, good_cases

CASES = good_cases(0.5)


def test_reference_main_test_cases(sample_fq):
    # main_test.go: TestOpenInputGzipByExtension / ByMagic write gzip.NewWriter output and expect the text back
    gz = gzip.compress(sample_fq)
    assert go.has_gzip_magic(gz) and go.gunzip(gz) == sample_fq
    # TestOpenInputPlain: no magic -> the input is passed through untouched
    assert not go.has_gzip_magic(sample_fq) and not go.has_gzip_magic(b"\x1f") and not go.has_gzip_magic(b"")


@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_python_gzip(name):
    gz, text = CASES[name]
    assert go.gunzip(gz) == text
    assert gzip.GzipFile(fileobj=io.BytesIO(gz)).read() == text


@pytest.mark.parametrize("name", sorted(bad_cases()))
def test_error_classes(name):
    gz, kind = bad_cases()[name]
    with pytest.raises(go.GunzipError) as e:
        go.gunzip(gz)
    if kind is not None:
        assert e.value.kind == kind
    # Python's reader skips zero padding behind a member, does not check the header CRC and takes names of any length; Go's does not
    if name not in ("trailing_zeros", "bad_hcrc", "name_too_long", "empty_input"):  # (and reads an empty input as an empty stream)
        with pytest.raises((OSError, EOFError, Exception)):
            gzip.GzipFile(fileobj=io.BytesIO(gz)).read()
