"""The file / stream front end of Image (image.d:859-1068, plugin.d:55-104) as far as it runs without a GPU: format from
the file name (the reference's own unittest, plugin.d:98-104, and its quirks), "Cannot open file", unidentified data,
detected formats without a loader, refusals of the save side. Loads and saves of real files are in the -m gpu tests."""
import io
import os

import pytest


@pytest.fixture()
def img(gb):
    from gamut_b200.image import Image
    return Image


def test_format_from_file_name(img):
    from gamut_b200.image import identifyImageFormatFromFilename as ident
    from gamut_b200.types import ImageFormat as F
    assert ident("mysueprduperphoto.jpg") == F.JPEG and ident("mysueprduperphoto.jfif") == F.JPEG           # plugin.d:100-103
    assert ident("c:\\compromising-photo.qoi") == F.QOI and ident("my/path/to/file.qoix") == F.QOIX
    assert ident("a.png") == F.PNG and ident("a.dib") == F.BMP and ident("a.tga") == F.TGA and ident("a.jif") == F.JPEG
    assert ident("a.PNG") == F.unknown and ident("a.") == F.unknown and ident("") == F.unknown and ident(None) == F.unknown
    assert ident("png") == F.PNG                                   # no '.': the whole name is the extension (plugin.d:64-69)
    assert ident("dir.d/file") == F.unknown and ident("archive.tar.gif") == F.GIF
    assert img.identifyFormatFromFileName("x.sqz") == F.SQZ


def test_load_errors_without_decoding(img, tmp_path):
    from gamut_b200.image import kStrCannotOpenFile, kStrImageFormatUnidentified, kStrImageFormatNoLoadSupport
    from gamut_b200.types import ImageFormat as F
    im = img()
    assert not im.loadFromFile(str(tmp_path / "missing.png")) and im.errorMessage() == kStrCannotOpenFile
    p = tmp_path / "noise.bin"
    p.write_bytes(b"nope, not an image")
    assert img.identifyFormatFromFile(str(p)) == F.unknown and img.identifyFormatFromFile(str(tmp_path / "missing")) == F.unknown
    assert not im.loadFromFile(str(p)) and im.errorMessage() == kStrImageFormatUnidentified
    g = tmp_path / "anim.bin"
    g.write_bytes(b"GIF89a" + b"\0" * 64)                          # detected by content; no loader in this build
    assert img.identifyFormatFromFile(str(g)) == F.GIF
    assert not im.loadFromFile(str(g)) and im.errorMessage() == kStrImageFormatNoLoadSupport
    d = tmp_path / "texture.dds"
    d.write_bytes(b"nope, not an image")                           # unknown content: the extension decides (image.d:866-870)
    assert not im.loadFromFile(str(d)) and im.errorMessage() == kStrImageFormatNoLoadSupport
    s = io.BytesIO(b"xx" + b"GIF87a" + b"\0" * 32)
    s.seek(2)
    assert img.identifyFormatFromStream(s) == F.GIF and s.tell() == 2
    assert not im.loadFromStream(s) and im.errorMessage() == kStrImageFormatNoLoadSupport
    assert not im.loadFromStream(io.BytesIO(b"")) and im.errorMessage() == kStrImageFormatUnidentified


def test_save_refusals_without_data(img, tmp_path):
    from gamut_b200.types import ImageFormat as F
    im = img()
    assert not im.saveToStream(F.QOI, io.BytesIO()) and not im.saveToStream(F.unknown, io.BytesIO())
    assert not im.saveToFile(str(tmp_path / "out.qoi")) and not im.saveToFile(F.TGA, str(tmp_path / "out.tga"))
    assert not im.saveToFile(str(tmp_path / "no-such-dir" / "out.qoi"))


def test_convert_cli_without_decoding(gb, tmp_path, capsys):
    """python -m gamut_b200.convert (examples/convert/source/main.d): option conflicts and load errors, the part of the CLI
    that runs without a GPU."""
    from gamut_b200 import convert
    for bad in (["a", "b", "--rgb", "--grey"], ["a", "b", "--alpha", "--drop-alpha"], ["a", "b", "-p", "--unpremul"]):
        with pytest.raises(SystemExit):
            convert.main(bad)
    assert convert.main([str(tmp_path / "missing.qoi"), str(tmp_path / "o.qoi")]) == 1
    assert "Cannot open file" in capsys.readouterr().err
    p = tmp_path / "noise.png"
    p.write_bytes(b"GIF89a" + b"\0" * 40)                          # content says GIF: no loader in this build
    assert convert.main([str(p), str(tmp_path / "o.tga")]) == 1
    assert "Cannot decode this image format in this build" in capsys.readouterr().err
