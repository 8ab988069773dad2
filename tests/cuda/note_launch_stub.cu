void note_launch(int) {}
