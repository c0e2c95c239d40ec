/* Embeds pyscf_b200/csrc/rys_tables.bin (made by tools/gen_rys_tables.py) into the library. */
__asm__(
    ".section .rodata\n"
    ".global b200jk_rys_blob\n"
    ".balign 16\n"
    "b200jk_rys_blob:\n"
    ".incbin \"" RYS_TABLE_PATH "\"\n"
    "b200jk_rys_blob_end:\n"
    ".global b200jk_rys_blob_size\n"
    ".balign 4\n"
    "b200jk_rys_blob_size:\n"
    ".int b200jk_rys_blob_end - b200jk_rys_blob\n"
    ".section .note.GNU-stack,\"\",@progbits\n");
