"""`create_video` of the reference (ffmpeg presentation video): nothing to do here."""


def create_video(output_dir, *a, **kw):
    return None
