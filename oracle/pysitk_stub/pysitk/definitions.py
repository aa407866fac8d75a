DIR_TMP = "/tmp"
